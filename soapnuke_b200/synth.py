"""Deterministic synthetic FASTQ generator and small FASTQ <-> SoA helpers.

Bases iid uniform ACGT; quality per position clip(round(N(mu_p, 6)), 2, 40) + 33 with mu_p linear
37 -> 28 over the read; per-pair class by u ~ U(0,1): 10 % adapter read-through at insert
k ~ U[20, L-10) (adapter written from k on both mates), 3 % N-rich (each base -> N w.p. 0.08),
5 % low-quality (all Q ~ U[2,8)), 4 % (polyg_frac) mate-2 polyG tail of length 5 + k/3, rest clean.
"""
import numpy as np

ADAPTER1 = b"AAGTCGGAGGCCAAGCGGTCTTAGGAAGACAA"            # -f, process_argv.cpp:1042
ADAPTER2 = b"AAGTCGGATCGTAGCCATGTCGTTCTGTGAGCCAAGGAGTTG"  # -r
SRNA_ADAPTER3 = b"TCGTATGCCGTCTTCTGCTTG"

SRNA_ADAPTER5 = b"GTTCAGAGTTCTACAGTCCGACGATC"

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def stride_for(max_len):
    return (max_len + 15) // 16 * 16


def _quals(rng, n, L):
    mu = np.linspace(37.0, 28.0, L, dtype=np.float32)
    q = rng.standard_normal((n, L), dtype=np.float32) * 6.0 + mu
    q = np.clip(np.rint(q), 2, 40).astype(np.uint8)
    return q


def gen_mate_arrays(rng, n, L):
    seq = _ACGT[rng.integers(0, 4, size=(n, L), dtype=np.uint8)]
    qual = _quals(rng, n, L)
    return seq, qual


def gen_pairs(n, L=150, seed=1002, polyg_frac=0.04, adapter1=ADAPTER1, adapter2=ADAPTER2,
              var_len=False, se=False, insert_range=None):
    """Return dict with seq1, qual1, len1 (and seq2, qual2, len2 unless se) as fixed-stride SoA
    (uint8 [n][stride], zero padded; uint16 [n])."""
    rng = np.random.default_rng(seed)
    stride = stride_for(L)
    out = {}
    u = rng.random(n)
    if insert_range is None:
        insert_range = (20, max(21, L - 10))
    k = rng.integers(insert_range[0], insert_range[1], size=n)
    cls_ada = u < 0.10
    cls_n = (u >= 0.10) & (u < 0.13)
    cls_lq = (u >= 0.13) & (u < 0.18)
    cls_pg = (u >= 0.18) & (u < 0.18 + polyg_frac)
    mates = (1,) if se else (1, 2)
    for m in mates:
        seq, qual = gen_mate_arrays(rng, n, L)
        ada = np.frombuffer(adapter1 if m == 1 else adapter2, dtype=np.uint8)
        if ada.size:
            idx = np.nonzero(cls_ada)[0]
            for i in idx:                      # 10 % of reads; small python loop is fine for tests
                s = min(int(k[i]), L); e = min(L, s + ada.size)
                seq[i, s:e] = ada[: e - s]
        nmask = cls_n[:, None] & (rng.random((n, L)) < 0.08)
        seq[nmask] = ord("N")
        lq = np.nonzero(cls_lq)[0]
        qual[lq] = rng.integers(2, 8, size=(lq.size, L), dtype=np.uint8)
        if m == 2 or se:
            pg = np.nonzero(cls_pg)[0]
            for i in pg:
                t = min(L, 5 + int(k[i]) // 3)
                seq[i, L - t:] = ord("G")
        length = np.full(n, L, dtype=np.uint16)
        if var_len:
            length = rng.integers(min(L, max(35, L // 2)), L + 1, size=n).astype(np.uint16)
        S = np.zeros((n, stride), dtype=np.uint8)
        Q = np.zeros((n, stride), dtype=np.uint8)
        S[:, :L] = seq
        Q[:, :L] = qual + 33
        if var_len:
            col = np.arange(stride)[None, :]
            pad = col >= length[:, None]
            S[pad] = 0
            Q[pad] = 0
        out[f"seq{m}"] = S
        out[f"qual{m}"] = Q
        out[f"len{m}"] = length
    out["n"] = n
    out["L"] = L
    out["stride"] = stride
    return out


def gen_srna(n, L=50, seed=1004, adapter5=SRNA_ADAPTER5, adapter3=SRNA_ADAPTER3, var_len=False):
    """SE small-RNA reads for the filtersRNA module: insert, then the 3' adapter (sometimes mutated), then
    random bases; classes without a 3' adapter, with an empty insert, with the 5' adapter's tail in front,
    with low qualities and with N's. Same dict layout as gen_pairs(se=True)."""
    rng = np.random.default_rng(seed)
    stride = stride_for(L)
    seq, qual = gen_mate_arrays(rng, n, L)
    a3 = np.frombuffer(adapter3, dtype=np.uint8)
    a5 = np.frombuffer(adapter5, dtype=np.uint8)
    u = rng.random(n)
    for i in range(n):
        if u[i] < 0.08:
            continue                                            # no 3' adapter
        if u[i] < 0.14:
            k = int(rng.integers(0, 4))                         # (nearly) empty insert
        else:
            k = int(rng.integers(15, min(36, L - 5)))
        piece = a3.copy()
        for _ in range(int(rng.integers(0, 3))):                # 0..2 substitutions
            piece[rng.integers(0, piece.size)] = _ACGT[rng.integers(0, 4)]
        e = min(L, k + piece.size)
        seq[i, k:e] = piece[: e - k]
        if 0.14 <= u[i] < 0.22:                                 # tail of the 5' adapter in front of the insert
            t = int(rng.integers(16, a5.size + 1))              # sRNA_hasAdapter wants >= adaRAr * adapter length matches
            t = min(t, L)
            seq[i, :t] = a5[-t:]
            if rng.random() < 0.5:
                seq[i, int(rng.integers(0, t))] = _ACGT[rng.integers(0, 4)]
        if 0.22 <= u[i] < 0.27:
            qual[i] = rng.integers(2, 8, size=L, dtype=np.uint8)
        if 0.27 <= u[i] < 0.32:
            seq[i, rng.integers(0, L, size=3)] = ord("N")
        if 0.32 <= u[i] < 0.36:
            seq[i, k - min(k, 12):k] = ord("G")                 # G-rich insert end (polyG after the adapter cut)
        if 0.36 <= u[i] < 0.40:
            j = int(rng.integers(0, L))
            seq[i, j] = seq[i, j] | 0x20                        # a lowercase base (never equals an uppercase adapter base)
    length = np.full(n, L, dtype=np.uint16)
    if var_len:
        length = rng.integers(max(8, L // 3), L + 1, size=n).astype(np.uint16)
    S = np.zeros((n, stride), dtype=np.uint8)
    Q = np.zeros((n, stride), dtype=np.uint8)
    S[:, :L] = seq
    Q[:, :L] = qual + 33
    col = np.arange(stride)[None, :]
    pad = col >= length[:, None]
    S[pad] = 0
    Q[pad] = 0
    return dict(seq1=S, qual1=Q, len1=length, n=n, L=L, stride=stride)


CONTAM1 = b"GATCGGAAGAGCACACGTCTGAACTCCAGTCAC"
CONTAM2 = b"CTGTCTCTTATACACATCTCCGAGCCCACGAGAC"
CONTAM3 = b"ACACTCTTTCCCTACACGACGCTCTTCCGATCT"


def revcomp(b):
    return bytes(b.translate(bytes.maketrans(b"ACGTN", b"TGCAN"))[::-1])


def add_contams(d, contams, seed, frac=0.15):
    """Plants (sometimes mutated / truncated / overhanging) copies of the contaminant sequences into a
    fraction of the reads of every mate of a gen_pairs() batch, plus a few N's next to them."""
    rng = np.random.default_rng(seed)
    n = d["n"]
    for m in (1, 2):
        if f"seq{m}" not in d:
            continue
        S, Ln = d[f"seq{m}"], d[f"len{m}"]
        for i in np.nonzero(rng.random(n) < frac)[0]:
            c = np.frombuffer(contams[int(rng.integers(0, len(contams)))], dtype=np.uint8).copy()
            l = int(Ln[i])
            for _ in range(int(rng.integers(0, 4))):
                c[rng.integers(0, c.size)] = _ACGT[rng.integers(0, 4)]
            mode = int(rng.integers(0, 4))
            if mode == 0:
                off = int(rng.integers(0, max(1, l - c.size + 1)))                 # inside
            elif mode == 1:
                off = -int(rng.integers(1, c.size - 5))                            # tail of the contaminant at the read's head
            elif mode == 2:
                off = l - int(rng.integers(5, c.size))                             # head of the contaminant at the read's tail
            else:
                c = c[: int(rng.integers(8, c.size))]; off = int(rng.integers(0, max(1, l - c.size + 1)))
            a0, b0 = max(off, 0), min(l, off + c.size)
            if b0 > a0:
                S[i, a0:b0] = c[a0 - off:b0 - off]
                if rng.random() < 0.3:
                    S[i, int(rng.integers(a0, b0))] = ord("N")
    return d


def read_ids(n, mate, first=0):
    return [b"@SYN:1:1101:%d:%d/%d" % ((first + i) // 1000, (first + i) % 1000, mate) for i in range(n)]


def tile_ids(n, mate, tiles=(1101, 1102, 1103, 1104, 2201)):
    """Old-style ids whose tile field (behind the 2nd ':') cycles irregularly through `tiles`."""
    return [b"@SYN:1:%d:%d:%d/%d" % (tiles[(i * 7 + i // 5) % len(tiles)], i // 1000, i % 1000, mate) for i in range(n)]


def fov_ids(n, mate):
    """Zebra-platform style ids with a CxxxRyyy field of view."""
    return [b"@V300012345L1C%03dR%03d%07d/%d" % (1 + (i * 3 + i // 7) % 4, 1 + (i // 3) % 5, i, mate) for i in range(n)]


def write_fastq(path, seq, qual, length, mate, first=0, gz=False, ids=None):
    import gzip
    n = seq.shape[0]
    if ids is None:
        ids = read_ids(n, mate, first)
    parts = []
    for i in range(n):
        l = int(length[i])
        parts.append(ids[i] + b"\n" + seq[i, :l].tobytes() + b"\n+\n" + qual[i, :l].tobytes() + b"\n")
    data = b"".join(parts)
    if gz:
        with gzip.open(path, "wb", compresslevel=2) as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)


def clean_fastq_bytes(seq, qual, length, results, mate, first=0, phred_shift=0, order=None, ids=None):
    """Rebuild the clean FASTQ text the reference writes (peprocess.cpp:3414) from per-read results.
    `order`: emission order of the read indices (abi.ref_output_order), default input order."""
    n = seq.shape[0]
    if ids is None:
        ids = read_ids(n, mate, first)
    parts = []
    for i in (range(n) if order is None else order):
        if results["category"][i] != 0:
            continue
        h = int(results["head_cut"][i]); l = int(results["clean_len"][i])
        q = qual[i, h:h + l]
        if phred_shift:
            q = (q.astype(np.int16) + phred_shift).astype(np.uint8)
        parts.append(ids[i] + b"\n" + seq[i, h:h + l].tobytes() + b"\n+\n" + q.tobytes() + b"\n")
    return b"".join(parts)


def parse_fastq(data, stride=None):
    """FASTQ text -> (ids, seq[n][stride], qual[n][stride], len[n])."""
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    n = len(lines) // 4
    ids = lines[0::4][:n]
    seqs = lines[1::4][:n]
    quals = lines[3::4][:n]
    mx = max((len(s) for s in seqs), default=0)
    if stride is None:
        stride = stride_for(max(mx, 1))
    S = np.zeros((n, stride), dtype=np.uint8)
    Q = np.zeros((n, stride), dtype=np.uint8)
    Ln = np.zeros(n, dtype=np.uint16)
    for i in range(n):
        l = len(seqs[i])
        S[i, :l] = np.frombuffer(seqs[i], dtype=np.uint8)
        Q[i, :l] = np.frombuffer(quals[i], dtype=np.uint8)
        Ln[i] = l
    return ids, S, Q, Ln


def write_fastq_fixed(path, seq, qual, L, mate, first=0):
    """Vectorised writer for uniform-length reads (bench-sized samples): fixed-width IDs
    `@SYN:1:1101:<7 digits>:<3 digits>/<mate>` so every record has the same byte length."""
    n = seq.shape[0]
    idx = np.arange(first, first + n, dtype=np.int64)
    hi, lo = idx // 1000, idx % 1000
    head = np.frombuffer(b"@SYN:1:1101:", dtype=np.uint8)
    idlen = head.size + 7 + 1 + 3 + 2
    rec = np.empty((n, idlen + 1 + L + 3 + L + 1), dtype=np.uint8)
    rec[:, :head.size] = head
    for k in range(7):
        rec[:, head.size + k] = (hi // 10 ** (6 - k)) % 10 + 48
    rec[:, head.size + 7] = ord(":")
    for k in range(3):
        rec[:, head.size + 8 + k] = (lo // 10 ** (2 - k)) % 10 + 48
    rec[:, head.size + 11] = ord("/")
    rec[:, head.size + 12] = 48 + mate
    o = idlen
    rec[:, o] = 10
    rec[:, o + 1:o + 1 + L] = seq[:, :L]
    rec[:, o + 1 + L] = 10
    rec[:, o + 2 + L] = ord("+")
    rec[:, o + 3 + L] = 10
    rec[:, o + 4 + L:o + 4 + 2 * L] = qual[:, :L]
    rec[:, o + 4 + 2 * L] = 10
    rec.tofile(path)

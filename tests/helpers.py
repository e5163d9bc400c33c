"""Shared test helpers: CPU replay harness loader, engine wrappers, comparison utilities."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from soapnuke_b200 import abi, synth  # noqa: E402

A1 = synth.ADAPTER1.decode()
A2 = synth.ADAPTER2.decode()
# BASELINE config 2 flags as CLI arguments and as make_params kwargs
CFG2_FLAGS = ["-f", A1, "-r", A2, "-J", "-l", "5", "-q", "0.5", "-n", "0.05", "-m", "15", "-p", "0.7",
              "-X", "50", "-g", "10", "-y", "20,30", "-x", "20,10"]
CFG2_KW = dict(adapter1=A1, adapter2=A2, ada_trim=True, low_qual=5, low_qual_ratio=0.5, n_ratio=0.05,
               mean_quality=15, highA_ratio=0.7, polyX_num=50, polyG_tail=10, trim_bad_tail=(20, 30),
               trim_bad_head=(20, 10))

_CT = None


def load_coretest():
    """Build + load tests/coretest/libcoretest.so (CPU replay of the kernel, test-only)."""
    global _CT
    if _CT is None:
        d = os.path.join(ROOT, "tests", "coretest")
        so = os.path.join(d, "libcoretest.so")
        deps = [os.path.join(d, "coretest.cpp")] + [os.path.join(ROOT, "soapnuke_b200", "csrc", f)
                                                    for f in ("filter_core.cuh", "filter_kernel.cuh", "dev_params.h", "text_core.cuh", "ws_core.cuh",
                                                              "ws_kernel.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
            extra = os.environ.get("SNK_CXXFLAGS", "").split()          # e.g. -DSNK_WS_J=2 (must match the engine build)
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wno-unknown-pragmas", "-fPIC", "-shared",
                                   "-I/usr/local/cuda/include", "-o", so, deps[0]] + extra)
        lib = C.CDLL(so)
        lib.coretest_filter.restype = C.c_int
        lib.coretest_filter.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Batch), C.POINTER(abi.Batch), C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.c_int, C.c_int, C.c_int]
        lib.coretest_filter_ws.restype = C.c_int
        lib.coretest_filter_ws.argtypes = lib.coretest_filter.argtypes
        lib.coretest_text_index_pack.restype = C.c_uint32
        lib.coretest_text_index_pack.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        lib.coretest_text_format.restype = C.c_uint64
        lib.coretest_text_format.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32] + \
            [C.c_int] * 6 + [C.c_uint32, C.c_void_p, C.c_void_p]
        _CT = lib
    return _CT


def core_replay(p, d, first=0, tile_r=0, grid=3, qb=-1):
    lib = load_coretest()
    n = d["seq1"].shape[0]
    b1 = abi.make_batch(d["seq1"], d["qual1"], d["len1"])
    r1 = np.zeros(n, dtype=abi.RESULT_DTYPE)
    r2 = np.zeros(n, dtype=abi.RESULT_DTYPE)
    st = np.zeros(p.n_slots * abi.SLOT_WORDS, dtype=np.uint64)
    err = C.c_uint32(0)
    if p.is_pe:
        b2 = abi.make_batch(d["seq2"], d["qual2"], d["len2"])
        lib.coretest_filter(C.byref(p), C.byref(b1), C.byref(b2), r1.ctypes.data, r2.ctypes.data, st.ctypes.data,
                            first, C.byref(err), tile_r, grid, qb)
    else:
        lib.coretest_filter(C.byref(p), C.byref(b1), None, r1.ctypes.data, None, st.ctypes.data, first,
                            C.byref(err), tile_r, grid, qb)
        r2 = None
    return r1, r2, st, err.value


def core_replay_ws(p, d, first=0, wpg=0, grid=3, qb=-1):
    """CPU replay of the warp-specialised kernel's work units (coretest_filter_ws); None when the shape is not served."""
    lib = load_coretest()
    n = d["seq1"].shape[0]
    st = np.zeros(p.n_slots * abi.SLOT_WORDS, dtype=np.uint64)
    err = C.c_uint32(0)
    b1 = abi.make_batch(d["seq1"], d["qual1"], d["len1"])
    r1 = np.zeros(n, dtype=abi.RESULT_DTYPE)
    if p.is_pe:
        r2 = np.zeros(n, dtype=abi.RESULT_DTYPE)
        b2 = abi.make_batch(d["seq2"], d["qual2"], d["len2"])
        rc = lib.coretest_filter_ws(C.byref(p), C.byref(b1), C.byref(b2), r1.ctypes.data, r2.ctypes.data, st.ctypes.data,
                                    first, C.byref(err), wpg, grid, qb)
    else:
        rc = lib.coretest_filter_ws(C.byref(p), C.byref(b1), None, r1.ctypes.data, None, st.ctypes.data, first,
                                    C.byref(err), wpg, grid, qb)
        r2 = None
    if rc:
        return None
    return r1, r2, st, err.value


def oracle_run(p, d, first=0, stats=None):
    import oracle_py as orc
    if p.is_pe:
        return orc.filter_pe(p, d, stats=stats, first_index=first)
    r1, st, err = orc.filter_se(p, d, stats=stats, first_index=first)
    return r1, None, st, err


class Engine:
    """Thin RAII wrapper over the C ABI (what a reference-side binding would do)."""

    def __init__(self, lib, params, device=0):
        self.lib = lib
        self.params = params
        self.h = C.c_void_p()
        rc = lib.snk_engine_create(C.byref(params), device, C.byref(self.h))
        if rc:
            raise RuntimeError(lib.snk_last_error().decode())

    def close(self):
        if self.h:
            self.lib.snk_engine_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, rc):
        if rc:
            raise RuntimeError(self.lib.snk_last_error().decode())

    def filter_host(self, d, first=0):
        n = d["seq1"].shape[0]
        b1 = abi.make_batch(d["seq1"], d["qual1"], d["len1"])
        r1 = np.zeros(n, dtype=abi.RESULT_DTYPE)
        if self.params.is_pe:
            b2 = abi.make_batch(d["seq2"], d["qual2"], d["len2"])
            r2 = np.zeros(n, dtype=abi.RESULT_DTYPE)
            self.check(self.lib.snk_filter_pe_host(self.h, C.byref(b1), C.byref(b2), r1.ctypes.data, r2.ctypes.data, first))
            return r1, r2
        self.check(self.lib.snk_filter_se_host(self.h, C.byref(b1), r1.ctypes.data, first))
        return r1, None

    def filter_text(self, texts, n, stride, first=0, lane=0, strip=1, pe_info=0, fasta=0, id_mode=0, fetch=True):
        """FASTQ text in -> (meta, [clean text per mate], [rec_off per mate], [results per mate]) through the text ABI."""
        fmt = abi.TextFormat(strip=strip, pe_info=pe_info, fasta=fasta, id_mode=id_mode)
        bufs = [np.frombuffer(t, dtype=np.uint8).copy() for t in texts]
        if self.params.is_pe:
            self.check(self.lib.snk_filter_pe_text_async(self.h, lane, bufs[0].ctypes.data, len(texts[0]), bufs[1].ctypes.data,
                                                         len(texts[1]), n, stride, C.byref(fmt), first))
        else:
            self.check(self.lib.snk_filter_se_text_async(self.h, lane, bufs[0].ctypes.data, len(texts[0]), n, stride, C.byref(fmt), first))
        meta = abi.TextMeta()
        self.check(self.lib.snk_text_meta_sync(self.h, lane, C.byref(meta)))
        if not fetch or meta.flags:
            return meta, None, None, None
        mates = 2 if self.params.is_pe else 1
        outs = [np.zeros(max(1, meta.out_bytes[m]), dtype=np.uint8) for m in range(mates)] + [None]
        offs = [np.zeros(n + 1, dtype=np.uint32) for m in range(mates)] + [None]
        ress = [np.zeros(n, dtype=abi.RESULT_DTYPE) for m in range(mates)] + [None]
        ptr = lambda a: a.ctypes.data if a is not None else None
        self.check(self.lib.snk_text_fetch_async(self.h, lane, ptr(outs[0]), ptr(outs[1]), ptr(offs[0]), ptr(offs[1]), ptr(ress[0]), ptr(ress[1])))
        self.check(self.lib.snk_engine_lane_sync(self.h, lane))
        return (meta, [outs[m][:meta.out_bytes[m]].tobytes() for m in range(mates)], offs[:mates], ress[:mates])

    def stats(self):
        st = np.zeros(self.params.n_slots * abi.SLOT_WORDS, dtype=np.uint64)
        self.check(self.lib.snk_engine_stats(self.h, st.ctypes.data))
        return st

    def error_flags(self):
        f = C.c_uint32(0)
        idx = C.c_uint64(0)
        self.check(self.lib.snk_engine_error_flags(self.h, C.byref(f), C.byref(idx)))
        return f.value, idx.value


def describe_stat_index(i):
    w = int(i) % abi.SLOT_WORDS
    slot = int(i) // abi.SLOT_WORDS
    if w < abi.FS_COUNT:
        return f"slot {slot} fs[{w}]"
    f = (w - abi.FS_COUNT) // abi.FILE_WORDS
    o = (w - abi.FS_COUNT) % abi.FILE_WORDS
    if o < abi.FILE_BS_OFF:
        return f"slot {slot} file {f} gs[{o}]"
    if o < abi.FILE_QS_OFF:
        o -= abi.FILE_BS_OFF
        return f"slot {slot} file {f} bs[pos {o // 5}][{o % 5}]"
    if o < abi.FILE_TS_OFF:
        o -= abi.FILE_QS_OFF
        return f"slot {slot} file {f} qs[pos {o // abi.QBINS}][q {o % abi.QBINS}]"
    o -= abi.FILE_TS_OFF
    return f"slot {slot} file {f} ts[{o // 1000}][{o % 1000}]"


def assert_same(got, want, what):
    """Bit-exact comparison of result records / statistics with a readable first difference."""
    r1, r2, st = got
    o1, o2, ost = want
    for nm, a, b in (("mate1", r1, o1), ("mate2", r2, o2)):
        if b is None:
            continue
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, f"{what}: {nm} results differ at {bad[:5]}: got {a[bad[:3]]} want {b[bad[:3]]}"
    bad = np.nonzero(st != ost)[0]
    assert bad.size == 0, (f"{what}: {bad.size} statistics words differ, first: " +
                           "; ".join(f"{describe_stat_index(i)} got {st[i]} want {ost[i]}" for i in bad[:5]))


# ---------------------------------------------------------------- FASTQ text path (test-side model)
def fastq_text(ids, seq, qual, length, eol=b"\n", last_newline=True):
    parts = []
    for i in range(len(ids)):
        l = int(length[i])
        parts.append(ids[i] + eol + seq[i, :l].tobytes() + eol + b"+" + eol + qual[i, :l].tobytes() + eol)
    data = b"".join(parts)
    if not last_newline and data.endswith(b"\n"):
        data = data[:-1]
    return data


def ref_id_transform(rid, mode):
    """read_filter.cpp:357-382"""
    if mode == 1:
        out = bytearray()
        cp = True
        for ch in rid:
            if ch == ord("#"):
                cp = False
            if cp:
                out.append(ch)
            elif ch == ord("/"):
                cp = True
                out.append(ch)
        return bytes(out)
    if mode == 2:
        k = rid.rfind(b":")
        return rid if k < 0 else rid[:k]
    return rid


def ref_clean_text(ids, seq, qual, res, mate, pe_info=0, fasta=0, id_mode=0, qshift=0):
    """peprocess.cpp:3383-3433 + :1617-1629 on per-read results; returns (text, rec_off[n+1])."""
    parts, off, total = [], [], 0
    for i in range(len(ids)):
        off.append(total)
        if res["category"][i] != 0:
            continue
        rid = ref_id_transform(ids[i], id_mode)
        rid += (b"/2" if mate else b"/1") * int(pe_info)
        h = int(res["head_cut"][i]); l = int(res["clean_len"][i])
        if fasta:
            rec = rid.replace(b"@", b">", 1) + b"\n" + seq[i, h:h + l].tobytes() + b"\n"
        else:
            q = (qual[i, h:h + l].astype(np.int16) + qshift).astype(np.uint8)
            rec = rid + b"\n" + seq[i, h:h + l].tobytes() + b"\n+\n" + q.tobytes() + b"\n"
        parts.append(rec)
        total += len(rec)
    off.append(total)
    return b"".join(parts), np.array(off, dtype=np.uint32)


def text_replay_index_pack(text, n, strip, stride):
    lib = load_coretest()
    buf = np.zeros(len(text) + 64, dtype=np.uint8)
    buf[:len(text)] = np.frombuffer(text, dtype=np.uint8)
    line_off = np.zeros(4 * n + 2, dtype=np.uint32)
    S = np.full((n, stride), 0xEE, dtype=np.uint8); Q = np.full((n, stride), 0xEE, dtype=np.uint8)
    Ln = np.zeros(n, dtype=np.uint16)
    mx = C.c_uint32(0)
    flags = lib.coretest_text_index_pack(buf.ctypes.data, len(text), n, strip, stride, line_off.ctypes.data, S.ctypes.data,
                                         Q.ctypes.data, Ln.ctypes.data, C.byref(mx))
    return flags, line_off[:4 * n + 1], S, Q, Ln, mx.value, buf


def text_replay_format(buf, nbytes, line_off, S, Q, res, mate, strip, pe_info=0, fasta=0, id_mode=0, qshift=0, lanes=32):
    lib = load_coretest()
    n = S.shape[0]
    out = np.zeros(nbytes + 2 * n + 64, dtype=np.uint8)
    rec_off = np.zeros(n + 1, dtype=np.uint32)
    res = np.ascontiguousarray(res)
    total = lib.coretest_text_format(buf.ctypes.data, line_off.ctypes.data, S.ctypes.data, Q.ctypes.data, res.ctypes.data, n,
                                     S.shape[1], mate, strip, pe_info, fasta, id_mode, qshift, lanes, out.ctypes.data,
                                     rec_off.ctypes.data)
    return out[:total].tobytes(), rec_off


def high_quality_mix(n=6000, L=100, seed=31, every=97):
    """PE batch where a few scattered records carry qualities above the shared-memory bins (Q42..Q60):
    most tiles take the unchecked histogram walk, some the checked one, and the two must add up."""
    d = synth.gen_pairs(n, L=L, seed=seed, var_len=True)
    rng = np.random.default_rng(seed)
    for key, lk in (("qual1", "len1"), ("qual2", "len2")):
        Q = d[key]
        for i in range(int(rng.integers(0, every)), n, every):
            l = int(d[lk][i])
            pos = rng.integers(0, l, size=max(1, l // 5))
            Q[i, pos] = 33 + rng.integers(42, 61, size=pos.size).astype(np.uint8)
    return d


def report_equal(ref_path, mine_path):
    """Byte equality of one report file, except for what the reference leaves to chance: in
    Distribution_of_Q20_Q30_bases_by_read_position_*.txt the raw columns of the rows behind the raw
    `read_length` (= length of the last record a worker saw) are printed from a `new float[]` that was
    never written (peprocess.cpp / seprocess.cpp:289-353), i.e. whatever the heap held - usually 0.0000,
    sometimes garbage. Those rows (where this writer prints 0.0000 for both raw columns) are compared
    on their other columns only."""
    import filecmp
    if filecmp.cmp(ref_path, mine_path, shallow=False):
        return True
    if "Distribution_of_Q20_Q30" not in os.path.basename(ref_path):
        return False
    a = open(ref_path).read().split("\n")
    b = open(mine_path).read().split("\n")
    if len(a) != len(b):
        return False
    for la, lb in zip(a, b):
        if la == lb:
            continue
        fa, fb = la.split("\t"), lb.split("\t")
        if len(fa) != len(fb) or len(fb) < 3 or fb[1] != "0.0000" or fb[2] != "0.0000":
            return False
        if fa[0] != fb[0] or fa[3:] != fb[3:]:
            return False
    return True

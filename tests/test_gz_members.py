"""CPU tier: the parallel multi-member gzip reader of the host driver (soapnuke_b200/host/gz_members.cpp) returns
exactly the byte stream zlib's gzread() / Python's gzip return for the same file - single member, many members of mixed
sizes, empty members, BGZF blocks (FEXTRA headers), gzip magic bytes inside member data (false candidates), trailing
garbage - for any thread count and read size, and reports truncated / corrupt streams as errors.
Reference behaviour it replaces: every worker inflates the whole file (peprocess.cpp:2014-2050, 2088-2131)."""
import gzip
import os
import random
import struct
import subprocess
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gzcat(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("gzcat") / "gzcat")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-o", out, os.path.join(ROOT, "tests", "gztest", "gzcat.cpp"),
                           os.path.join(ROOT, "soapnuke_b200", "host", "gz_members.cpp"), "-lz"])
    return out


def fastq_text(n, seed):
    rnd = random.Random(seed)
    return b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rnd.choice(b"ACGTN") for _ in range(70)), bytes(rnd.choice(b"#5<AFI") for _ in range(70)))
                    for i in range(n))


def bgzf(data):
    out = b""
    for i in range(0, len(data), 65280):
        blk = data[i:i + 65280]
        co = zlib.compressobj(2, zlib.DEFLATED, -15)
        raw = co.compress(blk) + co.flush()
        out += b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(raw) + 25) + raw + struct.pack("<II", zlib.crc32(blk), len(blk))
    return out + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def run(gzcat, path, threads, read_size):
    p = subprocess.run([gzcat, path, str(threads), str(read_size)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    info = dict(kv.split("=") for kv in p.stderr.decode().split() if "=" in kv)
    return p.returncode, p.stdout, info


TEXT = fastq_text(30000, 7)


def build_cases():
    rnd = random.Random(3)
    parts, pos = [], 0
    while pos < len(TEXT):
        n = rnd.choice([1, 100, 5000, 70000, 300000])
        parts.append(TEXT[pos:pos + n])
        pos += n
    multi = b"".join(gzip.compress(p, 2) for p in parts)
    evil = b"xx" + b"\x1f\x8b\x08\x00" * 50 + gzip.compress(b"a member stored inside a member must not be decoded") + TEXT[:100000]
    return {
        "single": (gzip.compress(TEXT, 2), TEXT, 1),
        "multi": (multi, TEXT, len(parts)),
        "empty_members_and_trailing_garbage": (gzip.compress(b"") + multi + gzip.compress(b"") + b"\0\0\0trailing garbage", TEXT, len(parts) + 2),
        "false_magic": (gzip.compress(evil, 0) + gzip.compress(TEXT[:5000], 2) + gzip.compress(evil, 0), evil + TEXT[:5000] + evil, 3),
        "bgzf": (bgzf(TEXT), TEXT, (len(TEXT) + 65279) // 65280 + 1),
    }


CASES = build_cases()


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("threads,read_size", [(1, 1 << 20), (4, 1 << 20), (8, 777), (3, 1 << 24)])
def test_member_reader_equals_gzread(gzcat, tmp_path, name, threads, read_size):
    blob, want, members = CASES[name]
    path = str(tmp_path / "in.gz")
    open(path, "wb").write(blob)
    rc, out, info = run(gzcat, path, threads, read_size)
    assert rc == 0
    assert out == want
    assert int(info["members"]) == members
    if name == "false_magic":
        assert int(info["cancelled"]) > 0          # the embedded headers were tried and dropped


def test_truncated_and_corrupt_streams_are_errors(gzcat, tmp_path):
    blob = CASES["multi"][0]
    path = str(tmp_path / "bad.gz")
    open(path, "wb").write(blob[:-100])
    rc, out, _ = run(gzcat, path, 4, 1 << 20)
    assert rc == 2 and TEXT.startswith(out)
    b2 = bytearray(blob)
    b2[len(b2) // 2] ^= 0x55
    open(path, "wb").write(bytes(b2))
    rc, out, _ = run(gzcat, path, 4, 1 << 20)
    assert rc == 2


def test_plain_text_is_left_to_gzread(gzcat, tmp_path):
    path = str(tmp_path / "plain.gz")
    open(path, "wb").write(TEXT[:4096])
    assert run(gzcat, path, 2, 4096)[0] == 3


# ---------------------------------------------------------------- the output side: soapnuke_b200/host/fast_deflate.cpp
@pytest.fixture(scope="module")
def fastgz(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fastgz") / "fastgz")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", out, os.path.join(ROOT, "tests", "gztest", "fastgz.cpp"),
                           os.path.join(ROOT, "soapnuke_b200", "host", "fast_deflate.cpp"), "-lz"])
    return out


def fast_deflate_cases():
    rnd = random.Random(11)
    return {
        "fastq": TEXT,
        "empty": b"", "one": b"A", "seven": b"ACGTACG", "eight": b"ACGTACGT", "nine": b"ACGTACGTA",
        "zeros": b"\0" * 300001,
        "random": bytes(rnd.randrange(256) for _ in range(300001)),
        "repeat_text": b"@SYN:1:1101:0000000:000/1\nACGT\n+\nIIII\n" * 20000,
        "period_251": bytes((i * 7) % 251 for i in range(200001)),
        "long_matches": bytes(rnd.randrange(256) for _ in range(1000)) * 300,
        "skewed": bytes((65 if rnd.random() < 0.999 else rnd.randrange(256)) for _ in range(200000)),
    }


FD_CASES = fast_deflate_cases()


@pytest.mark.parametrize("name", list(FD_CASES))
@pytest.mark.parametrize("piece", [4 << 20, 65537, 1001])
def test_fast_deflate_members_are_valid_gzip(fastgz, tmp_path, name, piece):
    """Every member the output encoder writes inflates (zlib, via Python's gzip: header, CRC-32 and ISIZE checked) to its
    input; the concatenation of members is what the clean .gz files consist of."""
    data = FD_CASES[name]
    path = str(tmp_path / "in.bin")
    open(path, "wb").write(data)
    p = subprocess.run([fastgz, path, str(piece)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert p.returncode == 0
    assert gzip.decompress(p.stdout) == data
    if name == "fastq":
        assert len(p.stdout) < 0.6 * len(data)

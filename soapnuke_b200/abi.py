"""ctypes mirror of include/snk_engine.h (the C ABI of the engine).

Only PODs and prototypes live here; the product code is the shared library built from
soapnuke_b200/csrc (CUDA) and soapnuke_b200/host (C++). There is deliberately NO CPU fallback:
if the library is missing, `load_engine()` raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ENGINE_LIB = os.path.join(HERE, "lib", "libsnk_engine.so")

ABI_VERSION = 5
MAX_READ_LEN = 1000
QBINS = 64
MAX_ADAPTERS = 8
MAX_ADAPTER_LEN = 128
MAX_SLOTS = 256
MAX_ID_FILTERS = 64
MAX_CONTAMS = 8
ID_FILTER_LEN = 8
LEN_MASK, PRE_TILE, PRE_FOV = 0x3FFF, 0x4000, 0x8000

FS_COUNT = 48
GS_COUNT = 16
FILE_COUNT = 4
TS_COUNT = 5
BS_WORDS = MAX_READ_LEN * 5
QS_WORDS = MAX_READ_LEN * QBINS
TS_WORDS = TS_COUNT * MAX_READ_LEN
FILE_WORDS = GS_COUNT + BS_WORDS + QS_WORDS + TS_WORDS
FILE_GS_OFF = 0
FILE_BS_OFF = GS_COUNT
FILE_QS_OFF = GS_COUNT + BS_WORDS
FILE_TS_OFF = GS_COUNT + BS_WORDS + QS_WORDS
SLOT_WORDS = FS_COUNT + FILE_COUNT * FILE_WORDS

RAW1, RAW2, CLEAN1, CLEAN2 = 0, 1, 2, 3
GS_READS, GS_BASES, GS_A, GS_C, GS_G, GS_T, GS_N, GS_Q20, GS_Q30, GS_LAST_KEY = range(10)

CATEGORY_NAMES = ["keep", "short", "long", "n", "highA", "polyX", "lowq", "meanq", "adapter", "empty", "no3adapter", "insertnull",
                  "tile", "fov", "contam", "gcontam"]
FS_BASE = {"adapter": 0, "n": 4, "highA": 8, "polyX": 12, "lowq": 16, "meanq": 20, "short": 24, "long": 28}


def slot_file_off(f):
    return FS_COUNT + f * FILE_WORDS


class Params(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("is_pe", C.c_int32),
        ("quality_phred", C.c_int32),
        ("out_quality_phred", C.c_int32),
        ("low_qual", C.c_int32),
        ("low_qual_ratio", C.c_float),
        ("mean_quality", C.c_int32),
        ("n_ratio", C.c_float),
        ("highA_ratio", C.c_float),
        ("polyG_tail", C.c_float),
        ("polyX_num", C.c_int32),
        ("min_read_length", C.c_int32),
        ("max_read_length", C.c_int32),
        ("ada_trim", C.c_int32),
        ("contam_trim", C.c_int32),
        ("ada_mis", C.c_int32 * 2),
        ("ada_mr", C.c_float * 2),
        ("ada_edge", C.c_int32 * 2),
        ("n_adapters", C.c_int32 * 2),
        ("adapter_len", (C.c_int32 * MAX_ADAPTERS) * 2),
        ("adapter", ((C.c_char * MAX_ADAPTER_LEN) * MAX_ADAPTERS) * 2),
        ("has_hard_trim", C.c_int32),
        ("hard_head", C.c_int32 * 2),
        ("hard_tail", C.c_int32 * 2),
        ("has_trim_bad_head", C.c_int32),
        ("bad_head_thr", C.c_int32),
        ("bad_head_max", C.c_int32),
        ("has_trim_bad_tail", C.c_int32),
        ("bad_tail_thr", C.c_int32),
        ("bad_tail_max", C.c_int32),
        ("index_remove", C.c_int32),
        ("max_base_quality", C.c_int32),
        ("n_slots", C.c_int32),
        ("slot_block", C.c_int64),
        ("srna", C.c_int32),
        ("ada_rctg", C.c_int32),
        ("ada_rar", C.c_float),
        ("ada_rma", C.c_int32),
        ("ada_rer", C.c_float),
        ("ada_rmm", C.c_int32),
        ("reserved", C.c_int32 * 2),
        ("seq_type1", C.c_int32),
        ("n_tile", C.c_int32),
        ("n_fov", C.c_int32),
        ("tile", (C.c_char * ID_FILTER_LEN) * MAX_ID_FILTERS),
        ("fov", (C.c_char * ID_FILTER_LEN) * MAX_ID_FILTERS),
        ("contam_discard", C.c_int32),
        ("n_contams", C.c_int32 * 2),
        ("contam_len", (C.c_int32 * MAX_CONTAMS) * 2),
        ("contam_seg_thr", (C.c_int32 * MAX_CONTAMS) * 2),
        ("contam", ((C.c_char * MAX_ADAPTER_LEN) * MAX_CONTAMS) * 2),
        ("n_gcontams", C.c_int32),
        ("gcontam_len", C.c_int32 * MAX_CONTAMS),
        ("gcontam_min_match", C.c_int32 * MAX_CONTAMS),
        ("gcontam_mismatch", C.c_int32 * MAX_CONTAMS),
        ("gcontam", (C.c_char * MAX_ADAPTER_LEN) * MAX_CONTAMS),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("seq", C.c_void_p),
        ("qual", C.c_void_p),
        ("len", C.c_void_p),
        ("n", C.c_uint32),
        ("stride", C.c_uint32),
    ]


class ReadResult(C.Structure):
    _fields_ = [
        ("head_cut", C.c_uint16),
        ("clean_len", C.c_uint16),
        ("category", C.c_uint8),
        ("mate_mask", C.c_uint8),
        ("adacut_pos", C.c_int16),
    ]


import numpy as _np

RESULT_DTYPE = _np.dtype([("head_cut", "<u2"), ("clean_len", "<u2"), ("category", "u1"),
                          ("mate_mask", "u1"), ("adacut_pos", "<i2")])
assert RESULT_DTYPE.itemsize == C.sizeof(ReadResult) == 8


def ref_threads_partition(threads_requested, nprocs=None, patch_size=None):
    """(n_slots, slot_block) exactly as the reference derives them:
    patchSize = T_requested*20000/8 (process_argv.cpp:541-544), T clamped to nprocs afterwards
    (process_argv.cpp:905-910), patch = 160/T (peprocess.cpp:81), block = patchSize*patch pairs
    (peprocess.cpp:2063: thread_read_block = 4*patchSize*patch lines)."""
    if nprocs is None:
        nprocs = os.cpu_count() or 1
    if not patch_size:                      # config key `patch=` overrides (process_argv.cpp:1370-1372)
        patch_size = threads_requested * 20000 // 8
    t = min(threads_requested, nprocs)
    patch = 160 // t
    return t, patch_size * patch, patch_size


def ref_output_order(n, threads_requested, nprocs=None, patch_size=None, gz_input=False, pe=True):
    """Order in which the reference emits the surviving records of an n-read input.

    Worker i owns blocks b with b % T == i (peprocess.cpp:2092) and writes each batch of patchSize
    reads to the temp file thread.<i>.<cycle>; the concat thread then appends the files cycle-major,
    thread-minor (peprocess.cpp:2957-2990). The cycle label of a full batch is computed from a line
    counter that, for plain-text input, has already advanced past the batch's last line
    (peprocess.cpp:2248 file1_line_num vs :2141 file2_line_num for .gz input). So with plain input
    the last batch before every cycle boundary is labelled with the NEXT cycle and is emitted after
    the other workers' next-cycle data. Inputs shorter than one cycle (block*T reads) come out in
    input order, and so does every SE run (seprocess.cpp:1080,1153 label before incrementing).
    Returns a list of read indices."""
    t, block, patch_size = ref_threads_partition(threads_requested, nprocs, patch_size)
    cyc = block * t
    keyed = []
    per_thread_fill = {}
    start = 0
    # walk the input block by block; inside a block batches are patch_size long
    for b0 in range(0, n, block):
        i = (b0 // block) % t
        b1 = min(n, b0 + block)
        for s0 in range(b0, b1, patch_size):
            s1 = min(b1, s0 + patch_size)
            full = (s1 - s0) == patch_size
            if full:
                label = (s1 // cyc) if (pe and not gz_input) else ((s1 - 1) // cyc)
            else:
                label = n // cyc            # flushed at EOF (peprocess.cpp:2166,2277)
            keyed.append((label, i, s0, s1))
    keyed.sort(key=lambda k: (k[0], k[1], k[2]))
    # Final concat pass (peprocess.cpp:2957-2966): in the LAST cycle the temp files are appended worker
    # by worker and the loop stops at the first worker that has no file for that cycle. A worker behind
    # such a gap loses its last-cycle file: with plain PE input that is the mislabelled batch before the
    # last cycle boundary whenever the input ends before worker T-2's block of the final cycle.
    # The reads are still counted in the clean statistics; they are just never written.
    last_label = max((k[0] for k in keyed), default=0)
    have = {k[1] for k in keyed if k[0] == last_label}
    first_gap = next((i for i in range(t) if i not in have), t)
    order = []
    for label, i, s0, s1 in keyed:
        if label == last_label and i > first_gap:
            continue
        order.extend(range(s0, s1))
    return order


def make_params(is_pe=True, adapter1=None, adapter2=None, ada_trim=False, low_qual=5, low_qual_ratio=0.5,
                mean_quality=-1, n_ratio=0.05, highA_ratio=-1.0, polyG_tail=-1.0, polyX_num=-1,
                min_read_length=30, max_read_length=-1, quality_phred=33, out_quality_phred=33,
                ada_mis=(2, 2), ada_mr=(0.5, 0.5), ada_edge=(6, 6), hard_trim=None,
                trim_bad_head=None, trim_bad_tail=None, threads=1, nprocs=None, max_base_quality=42,
                contam_trim=False, index_remove=False, patch_size=None, srna=False, ada_rctg=6, ada_rar=0.8,
                ada_rma=5, ada_rer=0.4, ada_rmm=4, tile=None, fov=None, seq_type1=False,
                contam1=None, contam2=None, ct_match_r="0.2",
                global_contams=None, glob_cotm_mR="", glob_cotm_mM=""):
    """Build snk_params the way process_argv.cpp would from `SOAPnuke filter` flags.
    Float thresholds go through double -> float exactly like `gp.x = atof(optarg)`."""
    p = Params()
    p.abi_version = ABI_VERSION
    p.is_pe = 1 if is_pe else 0
    p.quality_phred = quality_phred
    p.out_quality_phred = out_quality_phred
    p.low_qual = low_qual
    p.low_qual_ratio = low_qual_ratio
    p.mean_quality = mean_quality
    p.n_ratio = n_ratio
    p.highA_ratio = highA_ratio
    p.polyG_tail = polyG_tail
    p.polyX_num = polyX_num
    p.min_read_length = min_read_length
    p.max_read_length = max_read_length
    p.ada_trim = 1 if ada_trim else 0
    p.contam_trim = 1 if contam_trim else 0
    p.index_remove = 1 if index_remove else 0
    for m in range(2):
        p.ada_mis[m] = ada_mis[m]
        p.ada_mr[m] = ada_mr[m]
        p.ada_edge[m] = ada_edge[m]
    for m, ada in enumerate((adapter1, adapter2)):
        if ada is None:
            lst = []
        elif isinstance(ada, (str, bytes)):
            lst = [ada]
        else:
            lst = list(ada)
        assert len(lst) <= MAX_ADAPTERS
        p.n_adapters[m] = len(lst)
        for i, a in enumerate(lst):
            if isinstance(a, str):
                a = a.encode()
            assert len(a) < MAX_ADAPTER_LEN
            p.adapter_len[m][i] = len(a)
            p.adapter[m][i].value = a
    if hard_trim is not None:
        p.has_hard_trim = 1
        ht = list(hard_trim)
        if len(ht) == 2:
            ht = ht + [0, 0]
        p.hard_head[0], p.hard_tail[0], p.hard_head[1], p.hard_tail[1] = ht
    if trim_bad_head is not None:
        p.has_trim_bad_head = 1
        p.bad_head_thr, p.bad_head_max = trim_bad_head
    if trim_bad_tail is not None:
        p.has_trim_bad_tail = 1
        p.bad_tail_thr, p.bad_tail_max = trim_bad_tail
    p.max_base_quality = max_base_quality
    p.srna = 1 if srna else 0       # filtersRNA: adapter1 = 5' adapter, adapter2 = 3' adapter, SE only
    p.ada_rctg, p.ada_rar, p.ada_rma, p.ada_rer, p.ada_rmm = ada_rctg, ada_rar, ada_rma, ada_rer, ada_rmm
    # contam1 / contam2 / ctMatchR config values as strings, comma separated lists allowed (read_filter.cpp:188-206)
    p.contam_discard = 0 if contam_trim else 1
    import math
    for m, val in enumerate((contam1, contam2)):
        if not val:
            continue
        if "," not in val:
            seqs, thr = [val], [int(math.ceil(len(val) * float(ct_match_r)))]                        # double product (:609)
        else:
            seqs, mrs = val.split(","), ct_match_r.split(",")
            assert len(seqs) == len(mrs), "the number of ctMatchR value should equal to that of contam sequences"
            thr = [int(math.ceil(float(_np.float32(len(sq)) * _np.float32(float(mr))))) for sq, mr in zip(seqs, mrs)]   # float product (:514)
        assert len(seqs) <= MAX_CONTAMS
        p.n_contams[m] = len(seqs)
        for i, sq in enumerate(seqs):
            assert len(sq) < MAX_ADAPTER_LEN
            p.contam_len[m][i] = len(sq)
            p.contam_seg_thr[m][i] = thr[i]
            p.contam[m][i].value = sq.encode()
    if global_contams:                 # config keys global_contams / glob_cotm_mR / glob_cotm_mM (read_filter.cpp:927-944)
        seqs, mrs, mms = global_contams.split(","), glob_cotm_mR.split(","), glob_cotm_mM.split(",")
        assert len(seqs) == len(mrs) == len(mms) <= MAX_CONTAMS, "the number of global contamination sequences should equal to that of related parameters"
        p.n_gcontams = len(seqs)
        for i, sq in enumerate(seqs):
            p.gcontam_len[i] = len(sq)
            p.gcontam_min_match[i] = int(_np.float32(len(sq)) * _np.float32(float(mrs[i])))      # int(cl*min_matchRatio), float product
            p.gcontam_mismatch[i] = int(mms[i])
            p.gcontam[i].value = sq.encode()
    p.seq_type1 = 1 if seq_type1 else 0
    for name, val in (("tile", tile), ("fov", fov)):          # config keys tile= / fov= (comma separated)
        ents = [e for e in (val.split(",") if val else []) if 0 < len(e) <= ID_FILTER_LEN]
        assert len(ents) <= MAX_ID_FILTERS
        setattr(p, "n_" + name, len(ents))
        for i, e in enumerate(ents):
            getattr(p, name)[i].value = e.encode()
    n_slots, block, _ = ref_threads_partition(threads, nprocs, patch_size)
    p.n_slots = n_slots
    p.slot_block = block
    return p


def make_batch(seq, qual, length):
    """snk_batch over numpy arrays (kept alive by the caller)."""
    assert seq.dtype == _np.uint8 and qual.dtype == _np.uint8 and length.dtype == _np.uint16
    assert seq.flags.c_contiguous and qual.flags.c_contiguous and length.flags.c_contiguous
    n, stride = seq.shape
    assert qual.shape == (n, stride) and length.shape == (n,) and stride % 16 == 0
    b = Batch()
    b.seq = seq.ctypes.data
    b.qual = qual.ctypes.data
    b.len = length.ctypes.data
    b.n = n
    b.stride = stride
    return b


class TextFormat(C.Structure):
    _fields_ = [("strip", C.c_int32), ("pe_info", C.c_int32), ("fasta", C.c_int32), ("id_mode", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class TextMeta(C.Structure):
    _fields_ = [("out_bytes", C.c_uint64 * 2), ("kept", C.c_uint32), ("max_len", C.c_uint32), ("flags", C.c_uint32),
                ("bad_record", C.c_uint32)]


TEXT_STRIDE_OVERFLOW, TEXT_LEN_MISMATCH, TEXT_LINE_COUNT, TEXT_TOO_LONG = 1, 2, 4, 8

_PROTOS = {
    "snk_last_error": (C.c_char_p, []),
    "snk_abi_version": (C.c_int, []),
    "snk_stats_slot_words": (C.c_size_t, []),
    "snk_params_check": (C.c_int, [C.POINTER(Params)]),
    "snk_engine_create": (C.c_int, [C.POINTER(Params), C.c_int, C.POINTER(C.c_void_p)]),
    "snk_engine_destroy": (C.c_int, [C.c_void_p]),
    "snk_filter_pe_host": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_uint64]),
    "snk_filter_se_host": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_uint64]),
    "snk_engine_lanes": (C.c_int, [C.c_void_p]),
    "snk_filter_pe_async": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Batch), C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_uint64]),
    "snk_filter_se_async": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Batch), C.c_void_p, C.c_uint64]),
    "snk_engine_lane_sync": (C.c_int, [C.c_void_p, C.c_int]),
    "snk_filter_pe_text_async": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32,
                                           C.POINTER(TextFormat), C.c_uint64]),
    "snk_filter_se_text_async": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32,
                                           C.POINTER(TextFormat), C.c_uint64]),
    "snk_text_meta_sync": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(TextMeta)]),
    "snk_text_fetch_async": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snk_filter_pe_device": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "snk_filter_se_device": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_uint64, C.c_void_p]),
    "snk_engine_stats_reset": (C.c_int, [C.c_void_p]),
    "snk_engine_stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "snk_engine_stats_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "snk_engine_stats_to_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "snk_engine_stats_from_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "snk_engine_error_flags": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "snk_engine_launch_count": (C.c_uint64, [C.c_void_p]),
    "snk_engine_stage_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "snk_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "snk_host_free": (C.c_int, [C.c_void_p]),
    "snk_report_write_pe": (C.c_int, [C.POINTER(Params), C.c_void_p, C.c_char_p]),
    "snk_report_write_se": (C.c_int, [C.POINTER(Params), C.c_void_p, C.c_char_p]),
}
EXPORTED_SYMBOLS = sorted(_PROTOS)

_engine_lib = None


def load_engine():
    """dlopen the in-tree engine library. Fails loudly when it is missing: no fallback exists."""
    global _engine_lib
    if _engine_lib is None:
        if not os.path.exists(ENGINE_LIB):
            raise RuntimeError(
                f"{ENGINE_LIB} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the engine has no CPU fallback)")
        lib = C.CDLL(os.environ.get("SNK_ENGINE_LIB", ENGINE_LIB))       # SNK_ENGINE_LIB: a tuning build of the same engine
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.snk_abi_version() != ABI_VERSION:
            raise RuntimeError("ABI version mismatch between abi.py and libsnk_engine.so")
        _engine_lib = lib
    return _engine_lib

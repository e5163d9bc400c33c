/*
 * snk_engine.h — C ABI of the B200 FASTQ filter engine (drop-in for the SOAPnuke `filter` hot path).
 *
 * Boundary: every entry point below replaces one call the reference makes from
 * peProcess::thread_process_reads / seProcess::thread_process_reads (all paths are relative to the
 * reference tree):
 *
 *   reference call (file:line)                                   replaced by
 *   -----------------------------------------------------------  ---------------------------------------
 *   peProcess::filter_pe_fqs(PEcalOption*)  peprocess.cpp:1424   snk_filter_pe_host / snk_filter_pe_device
 *     -> C_pe_fastq_filter ctor (stat_read x2) sequence.cpp:182    (per-read counters, adapter_pos)
 *     -> pe_trim / fastq_trim              read_filter.cpp:338     (head/tail cuts, "longest cut wins")
 *     -> pe_discard                        sequence.cpp:198        (category + C_filter_stat counters)
 *   peProcess::stat_pe_fqs(opt,"raw")       peprocess.cpp:1076   same launch (raw tables)
 *   peProcess::stat_pe_fqs(opt,"clean")     peprocess.cpp:1961   same launch (clean tables)
 *   seProcess::filter_se_fqs(SEcalOption)   seprocess.cpp:871    snk_filter_se_host / snk_filter_se_device
 *   seProcess::stat_se_fqs(opt,...)         seprocess.cpp:632    same launch
 *   peProcess::merge_stat/update_stat       peprocess.cpp:1994,732  snk_engine_stats (+ snk_report_* on host)
 *   peProcess::print_stat                   peprocess.cpp:178    snk_report_write_pe
 *   seProcess::print_stat                   seprocess.cpp:96     snk_report_write_se
 *
 * Plain C types only: pointers, sizes, PODs. No torch / C++ types cross this boundary.
 * Every function returns 0 on success, non-zero on error (snk_last_error() gives the message);
 * the CLI layer prints "Error:..." and exit(1) to match the reference convention
 * (read_filter.cpp:250-253, 282-285; sequence.cpp:335-338).
 */
#ifndef SNK_ENGINE_H
#define SNK_ENGINE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNK_ABI_VERSION 5

/* ---- limits (global_variable.h:9-11 READ_MAX_LEN / MAX_QUAL) ---- */
#define SNK_MAX_READ_LEN   1000   /* READ_MAX_LEN: per-position tables have this many rows        */
#define SNK_QBINS          64     /* quality bins kept per position (reference: maxBaseQuality=42) */
#define SNK_MAX_ADAPTERS   8      /* adapters per mate (-f/-r list files), read_filter.cpp:177     */
#define SNK_MAX_ADAPTER_LEN 128
#define SNK_MAX_SLOTS      256    /* logical reference threads whose tables are kept apart         */
#define SNK_MAX_CONTAMS    8      /* contaminant sequences per mate (contam1 / contam2 lists)       */
#define SNK_MAX_ID_FILTERS 64     /* entries of the tile / fov removal lists                        */
#define SNK_ID_FILTER_LEN  8      /* a tile is up to 4 digits, a fov 8 characters (C001R003)        */

/* ---- parameters: the subset of C_global_parameter (global_parameter.h:20-190) the hot path reads ---- */
typedef struct snk_params {
    int32_t abi_version;          /* SNK_ABI_VERSION */
    int32_t is_pe;                /* 1 = peProcess, 0 = seProcess */
    int32_t quality_phred;        /* gp.qualityPhred (33|64) */
    int32_t out_quality_phred;    /* gp.outputQualityPhred */
    int32_t low_qual;             /* gp.lowQual (-l) */
    float   low_qual_ratio;       /* gp.lowQualityBaseRatio (-q); -1 disables */
    int32_t mean_quality;         /* gp.meanQuality (-m); -1 disables */
    float   n_ratio;              /* gp.n_ratio (-n); -1 disables */
    float   highA_ratio;          /* gp.highA_ratio (-p); -1 disables */
    float   polyG_tail;           /* gp.polyG_tail (-g); -1 disables */
    int32_t polyX_num;            /* gp.polyX_num (-X); -1 disables */
    int32_t min_read_length;      /* gp.min_read_length (-4); -1 disables */
    int32_t max_read_length;      /* gp.max_read_length; -1 disables */
    int32_t ada_trim;             /* gp.adapter_discard_or_trim == "trim" (-J) */
    int32_t contam_trim;          /* gp.contam_discard_or_trim == "trim" (only affects cut copy-back) */
    int32_t ada_mis[2];           /* adaMis, adaMis2 */
    float   ada_mr[2];            /* adaMR, adaMR2 */
    int32_t ada_edge[2];          /* adaEdge, adaEdge2 */
    int32_t n_adapters[2];        /* gp.ada1s.size(), gp.ada2s.size() */
    int32_t adapter_len[2][SNK_MAX_ADAPTERS];
    char    adapter[2][SNK_MAX_ADAPTERS][SNK_MAX_ADAPTER_LEN];
    int32_t has_hard_trim;        /* !gp.trim.empty() (-t) */
    int32_t hard_head[2];         /* head_trim_len per mate (peprocess.cpp:1692-1699) */
    int32_t hard_tail[2];
    int32_t has_trim_bad_head;    /* !gp.trimBadHead.empty() (-x "thr,maxlen") */
    int32_t bad_head_thr, bad_head_max;
    int32_t has_trim_bad_tail;    /* !gp.trimBadTail.empty() (-y "thr,maxlen") */
    int32_t bad_tail_thr, bad_tail_max;
    int32_t index_remove;         /* gp.index_remove: only turns trimming "on" (read_filter.cpp:348) */
    int32_t max_base_quality;     /* gp.maxBaseQuality (42) */
    /* logical-thread partition of the reference (peprocess.cpp:81,2063,2092): pair i belongs to
     * slot (i / slot_block) % n_slots. Tables are kept per slot so that update_stat's
     * partition-dependent merge (peprocess.cpp:732-1069) can be replayed on the host. */
    int32_t n_slots;              /* gp.threads_num after clamping; >=1 */
    int64_t slot_block;           /* gp.patchSize * patch ; >=1 */
    /* filtersRNA module (seProcess with the sRNA branches: read_filter.cpp:170-174, 432-438,
     * seprocess.cpp:899-902). adapter[0][0] is the 5' adapter (-f), adapter[1][0] the 3' adapter (-r).
     * Defaults global_parameter.h:54-58. */
    int32_t srna;                 /* gp.module_name == "filtersRNA" (SE only) */
    int32_t ada_rctg;             /* gp.adaRCtg: min 5' adapter continuous alignment length (6) */
    float   ada_rar;              /* gp.adaRAr: min alignment rate when finding the 5' adapter (0.8) */
    int32_t ada_rma;              /* gp.adaRMa: min alignment length when finding the 3' adapter (5) */
    float   ada_rer;              /* gp.adaREr: max error rate mismatch/match for the 3' adapter (0.4) */
    int32_t ada_rmm;              /* gp.adaRMm: max mismatches for the 3' adapter (4) */
    int32_t reserved[2];
    /* tile / fov removal lists (config keys `tile`, `fov`; check_tile_or_fov read_filter.cpp:14-79): a read
     * is dropped when the tile (fov) parsed from its id (stat_read, read_filter.cpp:86-148) equals a list
     * entry. IDs never cross the SoA boundary: callers of the SoA entry points mark such reads with
     * SNK_PRE_TILE / SNK_PRE_FOV in len[]; the FASTQ text entry points parse the ids on the device with
     * these lists. Entries are NUL padded. */
    int32_t seq_type1;            /* gp.seq_type == "1": the tile follows the 4th ':' instead of the 2nd */
    int32_t n_tile, n_fov;
    char    tile[SNK_MAX_ID_FILTERS][SNK_ID_FILTER_LEN];
    char    fov[SNK_MAX_ID_FILTERS][SNK_ID_FILTER_LEN];
    /* contaminant sequences (config keys contam1 / contam2 / ctMatchR; hasContam / hasContams,
     * read_filter.cpp:483-706): a read that contains one is dropped while contam_discard is set
     * (gp.contam_discard_or_trim == "discard", the default; `contam_trim` only switches that off).
     * contam_seg_thr = segMatchThr = (int)ceil(contamLen * ctMatchR), evaluated by the caller with the
     * reference's types: double product for a single contaminant (:609), float product for a list (:499,:514).
     * The matcher uses the mate's own adaMis / adaEdge (read 2: adaMis2 / adaEdge2, sequence.cpp:183-188). */
    int32_t contam_discard;
    int32_t n_contams[2];
    int32_t contam_len[2][SNK_MAX_CONTAMS];
    int32_t contam_seg_thr[2][SNK_MAX_CONTAMS];
    char    contam[2][SNK_MAX_CONTAMS][SNK_MAX_ADAPTER_LEN];
    /* global contaminants (config keys global_contams / glob_cotm_mR / glob_cotm_mM; hasGlobalContams /
     * global_contam_pos, read_filter.cpp:927-1053): every sequence is searched forward and reverse
     * complemented in both mates; a hit drops the read while contam_discard is set.
     * gcontam_min_match = int(len * matchRatio) with a float product (:970). */
    int32_t n_gcontams;
    int32_t gcontam_len[SNK_MAX_CONTAMS];
    int32_t gcontam_min_match[SNK_MAX_CONTAMS];
    int32_t gcontam_mismatch[SNK_MAX_CONTAMS];
    char    gcontam[SNK_MAX_CONTAMS][SNK_MAX_ADAPTER_LEN];
} snk_params;

/* ---- one mate of a batch, fixed-stride SoA ---- */
typedef struct snk_batch {
    const uint8_t*  seq;     /* [n][stride] bases, ASCII */
    const uint8_t*  qual;    /* [n][stride] qualities, ASCII */
    const uint16_t* len;     /* [n] read lengths (1..stride) in bits 0-13; bits 14/15 = SNK_PRE_TILE / SNK_PRE_FOV */
    uint32_t        n;       /* reads in this batch */
    uint32_t        stride;  /* bytes per row, multiple of 16, <= 1008 */
} snk_batch;

/* len[] flag bits: the read's id selected it for removal by the tile / fov lists (for a pair only mate 1's
 * flags count, sequence.cpp:213-230) */
#define SNK_LEN_MASK  0x3FFFu
#define SNK_PRE_TILE  0x4000u
#define SNK_PRE_FOV   0x8000u

/* ---- per-read result record (8 bytes) ---- */
/* category codes, in pe_discard / se_discard priority order (sequence.cpp:198-387, 76-178) */
enum snk_category {
    SNK_KEEP = 0,
    SNK_DROP_SHORT = 1,          /* "Reads too short" */
    SNK_DROP_LONG = 2,           /* "Reads too long" */
    SNK_DROP_N = 3,              /* "Reads with n rate exceed" */
    SNK_DROP_HIGHA = 4,          /* "Reads with highA" */
    SNK_DROP_POLYX = 5,          /* "Reads with polyX" */
    SNK_DROP_LOWQ = 6,           /* "Reads with low quality" */
    SNK_DROP_MEANQ = 7,          /* "Reads with low mean quality" */
    SNK_DROP_ADAPTER = 8,        /* "Reads with adapter" */
    SNK_DROP_EMPTY = 9,          /* min_read_length==-1 and a mate was emptied (sequence.cpp:245-249), uncounted */
    SNK_DROP_NO3ADAPTER = 10,    /* filtersRNA: no 3' adapter found (sequence.cpp:36-39), counted but never reported */
    SNK_DROP_INSERTNULL = 11,    /* filtersRNA: 3' adapter within the first 3 bases (sequence.cpp:40-44), never reported */
    SNK_DROP_TILE = 12,          /* "Reads with filtered tile" (first test of pe_discard / se_discard) */
    SNK_DROP_FOV = 13,           /* "Reads with filtered fov" */
    SNK_DROP_CONTAM = 14,        /* "Reads with contam sequence" */
    SNK_DROP_GCONTAM = 15        /* "Reads with global contam sequence" */
};
typedef struct snk_read_result {
    uint16_t head_cut;     /* bases removed from the 5' end of this mate */
    uint16_t clean_len;    /* length after trimming (0 when emptied) */
    uint8_t  category;     /* snk_category of the pair (same value on both mates) */
    uint8_t  mate_mask;    /* pe_dis(): 1 = fq1 triggered, 2 = fq2, 3 = both (0 when kept) */
    int16_t  adacut_pos;   /* C_fastq::adacut_pos: len - adapter_pos, or -1 when no adapter */
} snk_read_result;

/* ---- statistics: flat uint64 tables, one block per slot ---- */
/* C_filter_stat counters (global_variable.h:66-86), index into the fs[] block */
enum snk_fs {
    SNK_FS_ADAPTER = 0, SNK_FS_ADAPTER1, SNK_FS_ADAPTER2, SNK_FS_ADAPTER_OV,
    SNK_FS_N, SNK_FS_N1, SNK_FS_N2, SNK_FS_N_OV,
    SNK_FS_HIGHA, SNK_FS_HIGHA1, SNK_FS_HIGHA2, SNK_FS_HIGHA_OV,
    SNK_FS_POLYX, SNK_FS_POLYX1, SNK_FS_POLYX2, SNK_FS_POLYX_OV,
    SNK_FS_LOWQ, SNK_FS_LOWQ1, SNK_FS_LOWQ2, SNK_FS_LOWQ_OV,
    SNK_FS_MEANQ, SNK_FS_MEANQ1, SNK_FS_MEANQ2, SNK_FS_MEANQ_OV,
    SNK_FS_SHORT, SNK_FS_SHORT1, SNK_FS_SHORT2, SNK_FS_SHORT_OV,
    SNK_FS_LONG, SNK_FS_LONG1, SNK_FS_LONG2, SNK_FS_LONG_OV,
    SNK_FS_NO3ADAPTER,           /* fs.no_3_adapter_num (filtersRNA) */
    SNK_FS_INSERTNULL,           /* fs.int_insertNull_num (filtersRNA) */
    SNK_FS_TILE,                 /* fs.tile_num */
    SNK_FS_FOV,                  /* fs.fov_num */
    SNK_FS_CONTAM, SNK_FS_CONTAM1, SNK_FS_CONTAM2, SNK_FS_CONTAM_OV,   /* fs.include_contam_seq_num[1|2|_overlap] */
    SNK_FS_GCONTAM, SNK_FS_GCONTAM1, SNK_FS_GCONTAM2, SNK_FS_GCONTAM_OV,   /* fs.include_global_contam_seq_num[1|2|_overlap] */
    SNK_FS_COUNT = 48
};
/* C_general_stat (global_variable.h:88-100), index into a file block's gs[] */
enum snk_gs {
    SNK_GS_READS = 0, SNK_GS_BASES, SNK_GS_A, SNK_GS_C, SNK_GS_G, SNK_GS_T, SNK_GS_N,
    SNK_GS_Q20, SNK_GS_Q30,
    SNK_GS_LAST_KEY,   /* max over reads of ((global_index+1)<<16 | length): gives gs.read_length
                          = length of the LAST record this slot saw (peprocess.cpp:1202,1419) */
    SNK_GS_COUNT = 16
};
/* which FASTQ set a file block describes (C_global_variable, global_variable.h:136-143) */
enum snk_file { SNK_RAW1 = 0, SNK_RAW2 = 1, SNK_CLEAN1 = 2, SNK_CLEAN2 = 3, SNK_FILE_COUNT = 4 };
/* C_reads_trim_stat member order (global_variable.h:118-124): hlq, ht, ta, tlq, tt are contiguous
 * arrays of READ_MAX_LEN; the reference indexes ta/tlq/tt with possibly NEGATIVE indices
 * (raw_length==0 on raw fq1 records, peprocess.cpp:1124-1140), which land in the preceding array.
 * The flat ts[] block keeps that layout so the spill is reproduced bit for bit. */
enum snk_ts { SNK_TS_HLQ = 0, SNK_TS_HT = 1, SNK_TS_TA = 2, SNK_TS_TLQ = 3, SNK_TS_TT = 4, SNK_TS_COUNT = 5 };

#define SNK_BS_WORDS  (SNK_MAX_READ_LEN * 5)                 /* position_acgt_content[pos][ACGTN]  */
#define SNK_QS_WORDS  (SNK_MAX_READ_LEN * SNK_QBINS)         /* position_qual[pos][q]              */
#define SNK_TS_WORDS  (SNK_TS_COUNT * SNK_MAX_READ_LEN)
#define SNK_FILE_WORDS (SNK_GS_COUNT + SNK_BS_WORDS + SNK_QS_WORDS + SNK_TS_WORDS)
#define SNK_FILE_GS_OFF 0
#define SNK_FILE_BS_OFF (SNK_GS_COUNT)
#define SNK_FILE_QS_OFF (SNK_GS_COUNT + SNK_BS_WORDS)
#define SNK_FILE_TS_OFF (SNK_GS_COUNT + SNK_BS_WORDS + SNK_QS_WORDS)
#define SNK_SLOT_WORDS (SNK_FS_COUNT + SNK_FILE_COUNT * SNK_FILE_WORDS)
#define SNK_SLOT_FILE_OFF(f) (SNK_FS_COUNT + (size_t)(f) * SNK_FILE_WORDS)

/* ---- engine ---- */
typedef struct snk_engine snk_engine;

const char* snk_last_error(void);
int  snk_abi_version(void);
/* words (uint64) of one slot's statistics block == SNK_SLOT_WORDS */
size_t snk_stats_slot_words(void);

/* Validate parameters the way the hot path needs them (adapter length vs adaMis: the reference
 * divides by (adptLen-5)/(adaMis+1), read_filter.cpp:714-715). 0 = ok. */
int snk_params_check(const snk_params* p);

int snk_engine_create(const snk_params* p, int device, snk_engine** out);
int snk_engine_destroy(snk_engine* e);

/* Host-buffer entry points: replace filter_pe_fqs + stat_pe_fqs(raw) + stat_pe_fqs(clean) for one
 * batch. Copies the batch host->device, runs the kernels, copies the per-read results back into
 * out1/out2 (host) and accumulates the statistics on the device. Synchronous on return.
 * first_index = number of pairs (reads for SE) that precede this batch in the input; it selects
 * the slot per read. These two calls may be issued from several threads at once (the reference calls
 * filter_pe_fqs from its T workers, peprocess.cpp:1915): they take turns on the engine's first lane. The
 * asynchronous entry points below are single-producer per lane. */
int snk_filter_pe_host(snk_engine* e, const snk_batch* r1, const snk_batch* r2,
                       snk_read_result* out1, snk_read_result* out2, uint64_t first_index);
int snk_filter_se_host(snk_engine* e, const snk_batch* r1, snk_read_result* out1, uint64_t first_index);

/* Asynchronous variant of the above for the pinned-buffer driver: enqueue on one of the engine's
 * `lanes` (copy-in, kernel, copy-out on that lane's stream) and return; snk_engine_lane_sync waits.
 * Buffers must stay valid (and should be pinned) until the lane is synchronised. */
int snk_engine_lanes(snk_engine* e);
int snk_filter_pe_async(snk_engine* e, int lane, const snk_batch* r1, const snk_batch* r2,
                        snk_read_result* out1, snk_read_result* out2, uint64_t first_index);
int snk_filter_se_async(snk_engine* e, int lane, const snk_batch* r1, snk_read_result* out1, uint64_t first_index);
int snk_engine_lane_sync(snk_engine* e, int lane);

/* ---- FASTQ text entry points (the steps right before and right after the path) ----
 *
 *   reference (file:line)                                            replaced by
 *   ---------------------------------------------------------------  --------------------------------------
 *   sub_thread's line loop + C_fastq fill   peprocess.cpp:2090-2131   line index + row packing kernels
 *     (.gz: every line loses spaceNum chars), :2198-2239 (plain: 1)     (snk_text_format.strip)
 *   fastq_trim index removal                read_filter.cpp:357-382   id_mode
 *   preOutput (/1 /2)                       peprocess.cpp:1617-1629   pe_info
 *   output_fastqs                           peprocess.cpp:3383-3433   format kernel: the clean records of the
 *                                           seprocess.cpp:2302-2352     batch as one contiguous text per mate
 *
 * The host hands over the raw text of n_records whole records per mate (4 lines each; the last line
 * of the input may lack its newline) and gets back the clean FASTQ/FASTA text, in input order, plus
 * the byte offset of every record in it. Statistics accumulate on the device exactly as with the
 * SoA entry points. Sequence: snk_filter_*_text_async -> snk_text_meta_sync (sizes, flags) ->
 * snk_text_fetch_async into buffers of at least out_bytes[m] -> snk_engine_lane_sync. */
typedef struct snk_text_format {
    int32_t strip;        /* characters every input line loses at its end, newline included: 1 for plain
                             input, spaceNum of the first line for .gz input (peprocess.cpp:2066-2076) */
    int32_t pe_info;      /* how many "/1" ("/2") suffixes preOutput appends to the id: gp.whether_add_pe_info (0|1); 2 for the
                             clean records when trim files are written as well (preOutput runs twice on them) */
    int32_t fasta;        /* gp.output_file_type == "fasta" */
    int32_t id_mode;      /* 0: ids unchanged; 1: gp.index_remove with seqType "0"; 2: gp.index_remove otherwise */
    int32_t reserved[4];
} snk_text_format;
enum snk_text_flags {
    SNK_TEXT_STRIDE_OVERFLOW = 1,  /* a read is longer than `stride`: nothing was filtered or counted; resubmit with
                                      stride >= roundup16(max_len) */
    SNK_TEXT_LEN_MISMATCH = 2,     /* sequence and quality of record bad_record differ in length */
    SNK_TEXT_LINE_COUNT = 4,       /* the text does not hold 4 * n_records lines */
    SNK_TEXT_TOO_LONG = 8          /* a read exceeds SNK_MAX_READ_LEN */
};
typedef struct snk_text_meta {
    uint64_t out_bytes[2];  /* size of the clean text per mate */
    uint32_t kept;          /* records (pairs) kept */
    uint32_t max_len;       /* longest read of the batch */
    uint32_t flags;         /* snk_text_flags */
    uint32_t bad_record;    /* first offending record for LEN_MISMATCH / TOO_LONG */
} snk_text_meta;
/* text must be pinned for the copy to overlap; at most 3.75 GiB per mate and call */
int snk_filter_pe_text_async(snk_engine* e, int lane, const char* text1, size_t bytes1, const char* text2, size_t bytes2,
                             uint32_t n_records, uint32_t stride, const snk_text_format* fmt, uint64_t first_index);
int snk_filter_se_text_async(snk_engine* e, int lane, const char* text1, size_t bytes1, uint32_t n_records, uint32_t stride,
                             const snk_text_format* fmt, uint64_t first_index);
/* waits for the lane, then reports sizes and flags of its last text submission */
int snk_text_meta_sync(snk_engine* e, int lane, snk_text_meta* out);
/* enqueue the copies back to the host (any pointer may be NULL): clean text, rec_off[n_records + 1]
 * (byte offset of every record in the clean text; equal neighbours = record dropped), result records */
int snk_text_fetch_async(snk_engine* e, int lane, char* out1, char* out2, uint32_t* rec_off1, uint32_t* rec_off2,
                         snk_read_result* res1, snk_read_result* res2);

/* Device-resident entry points: all pointers are DEVICE pointers (e.g. torch tensors' data_ptr()),
 * `stream` is a cudaStream_t (0 = legacy default stream). Asynchronous. */
int snk_filter_pe_device(snk_engine* e, const snk_batch* d_r1, const snk_batch* d_r2,
                         snk_read_result* d_out1, snk_read_result* d_out2,
                         uint64_t first_index, void* stream);
int snk_filter_se_device(snk_engine* e, const snk_batch* d_r1, snk_read_result* d_out1,
                         uint64_t first_index, void* stream);

/* Statistics. The device keeps n_slots blocks of SNK_SLOT_WORDS uint64. */
int snk_engine_stats_reset(snk_engine* e);
/* copy all slots to host: dst has n_slots * SNK_SLOT_WORDS uint64 */
int snk_engine_stats(snk_engine* e, uint64_t* dst);
/* device pointer of the live table (for NCCL all-reduce by the caller) and its size in words */
int snk_engine_stats_device(snk_engine* e, uint64_t** d_ptr, size_t* words);
/* device-to-device copies of the whole table out of / into the engine (e.g. to a torch tensor that
 * is then all-reduced with NCCL by the caller). Asynchronous on `stream`. */
int snk_engine_stats_to_device(snk_engine* e, void* d_dst, void* stream);
int snk_engine_stats_from_device(snk_engine* e, const void* d_src, void* stream);
/* sticky error flags raised by kernels: bit0 = unrecognized base (read_filter.cpp:282),
 * bit1 = quality outside [0,SNK_QBINS), bit2 = low quality ratio > 1 (sequence.cpp:335),
 * bit3 = a len[] entry exceeds the batch stride or SNK_MAX_READ_LEN (SoA entry points; the row is not processed) */
int snk_engine_error_flags(snk_engine* e, uint32_t* flags, uint64_t* first_bad_index);
/* number of kernel launches issued by this engine so far */
uint64_t snk_engine_launch_count(snk_engine* e);
/* Device time per stage of the FASTQ text path, in milliseconds, summed over every batch of every lane whose work has
 * been synchronised (CUDA events on the lane streams). The reference only logs wall-clock lines per 5 s poll
 * (peprocess.cpp:3039); this is the per-stage view SURVEY.md section 5 asks for. The same stages are NVTX ranges
 * ("snk:text_submit", "snk:text_fetch", "snk:lane_sync") for timeline tools. ms has SNK_STAGE_COUNT entries. */
enum snk_stage { SNK_STAGE_H2D = 0, SNK_STAGE_INDEX_PACK, SNK_STAGE_FILTER, SNK_STAGE_FORMAT, SNK_STAGE_D2H, SNK_STAGE_COUNT };
int snk_engine_stage_times(snk_engine* e, double* ms);
/* pinned host memory helpers for callers without a CUDA runtime binding */
int snk_host_alloc(void** p, size_t bytes);
int snk_host_free(void* p);

/* ---- host-side report writer (print_stat + update_stat), no GPU needed ---- */
/* stats = n_slots blocks, as returned by snk_engine_stats (or summed across GPUs slot by slot).
 * Writes the 10 (PE) / 6 (SE) report files of peprocess.cpp:178-731 / seprocess.cpp:96-434 into out_dir. */
int snk_report_write_pe(const snk_params* p, const uint64_t* stats, const char* out_dir);
int snk_report_write_se(const snk_params* p, const uint64_t* stats, const char* out_dir);

#ifdef __cplusplus
}
#endif
#endif /* SNK_ENGINE_H */

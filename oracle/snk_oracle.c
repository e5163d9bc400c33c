/*
 * snk_oracle.c — TEST INFRASTRUCTURE ONLY. CPU restatement (plain C) of the reference's per-read
 * filter / trim / statistics path, used as the checker for the CUDA engine. Nothing in the product
 * path (soapnuke_b200/) may include, link or call this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the UNMODIFIED reference binary compiled into oracle/_ref/SOAPnuke
 * (oracle/Makefile): tests/test_oracle.py runs both on the same FASTQ and compares the
 * clean FASTQ and all report files byte for byte, and tests/golden/ holds outputs of that binary.
 *
 * Each function cites the reference file:line (relative to the reference tree) it follows.
 * It works on the same fixed-stride SoA batches and fills the same result/statistics layouts as the
 * engine (include/snk_engine.h supplies only those POD definitions).
 */
#define _GNU_SOURCE
#include <math.h>
#include <limits.h>
#include <string.h>
#include <stdio.h>
#include "../include/snk_engine.h"

#define ORC_EXPORT __attribute__((visibility("default")))

static int float_to_int_x86(float f)
{
    /* the reference assigns float quotients to int (read_filter.cpp:724,769); on x86-64 cvttss2si
       yields INT_MIN for NaN / out-of-range, which is what the compiled reference does. */
    if (!(f == f) || f >= 2147483648.0f || f < -2147483648.0f) return INT_MIN;
    return (int)f;
}

/* read_filter.cpp:707-790 adapter_pos(). Out-of-range read positions (reads shorter than the
 * adapter window; UB in the reference) compare as mismatches here. */
ORC_EXPORT int orc_adapter_pos(const uint8_t* read, int readLen, const uint8_t* adapter, int adptLen,
                               int adaMis, float adaMR, int adaEdge)
{
    if (adptLen == 0) return -1;
    const int minEdge5 = 5;
    float misGrad5 = (float)((adptLen - minEdge5) / (adaMis + 1));   /* :714 integer division first */
    float misGrad = (float)((adptLen - adaEdge) / (adaMis + 1));     /* :715 */
    int segMatchThr = (int)ceilf((float)adptLen * adaMR);            /* :717 */
    int r1, mis, seg, budget;
    for (r1 = 1; r1 <= minEdge5; ++r1) {                             /* :720-742 phase 1 */
        mis = 0; seg = 0;
        budget = float_to_int_x86((float)(adptLen - r1) / misGrad5);
        for (int c = 0; c < adptLen - r1; ++c) {
            int same = (c < readLen) && adapter[r1 + c] == read[c];
            if (same) { if (++seg >= segMatchThr) return 0; }
            else { mis++; seg = 0; if (mis > budget) break; }
        }
        if (mis <= budget) return 0;
    }
    for (r1 = 0; r1 <= readLen - adptLen; ++r1) {                    /* :743-764 phase 2 */
        seg = 0; mis = 0;
        for (int c = 0; c < adptLen; ++c) {
            if (adapter[c] == read[r1 + c]) { if (++seg >= segMatchThr) return r1; }
            else { mis++; seg = 0; if (mis > adaMis) break; }
        }
        if (mis <= adaMis) return r1;
    }
    for (r1 = 0; r1 < adptLen - adaEdge; ++r1) {                     /* :765-788 phase 3 */
        mis = 0; seg = 0;
        budget = float_to_int_x86((float)r1 / misGrad);
        int base = readLen - r1 - adaEdge;
        for (int c = 0; c < r1 + adaEdge; ++c) {
            int idx = base + c;
            int same = (idx >= 0 && idx < readLen) && adapter[c] == read[idx];
            if (same) { if (++seg >= segMatchThr) return base; }
            else { mis++; seg = 0; if (mis > budget) break; }
        }
        if (mis <= budget) return base;
    }
    return -1;
}

/* read_filter.cpp:791-862 sRNA_findAdapter(): ungapped alignments of the 3' adapter against the read,
 * first with the adapter's offsets 2,1,0 at read position 0, then adapter offset 0 at read positions
 * 1..readLen-adaRMa. 'N' in the read is neither a match nor a mismatch. Returns the read position of the
 * preferred accepted alignment (a later one replaces the current one only if it has no more mismatches
 * and no fewer matches) or -1. */
ORC_EXPORT int orc_srna_find_adapter(const uint8_t* read, int readLen, const uint8_t* adapter, int adptLen,
                                     int adaRMa, int adaRMm, float adaREr)
{
    int startPos = -1;
    if (adptLen == 0) return -1;
    int a1 = 2, flagType = 0, misTmp = 0, totalMapTmp = 0;
    for (int r1 = 0; r1 <= readLen - adaRMa;) {
        int len1 = adptLen - a1, len2 = readLen - r1;
        int len = len1 < len2 ? len1 : len2;
        int mis = 0, totalMap = 0;
        for (int c = 0; c < len; c++) {
            if (read[r1 + c] == 'N') continue;
            if (adapter[a1 + c] == read[r1 + c]) totalMap++;
            else mis++;
        }
        int misAndMap = mis + totalMap;
        float rate = 1.0 * mis / totalMap;                 /* double division, then float (:832) */
        if (mis <= adaRMm && misAndMap >= adaRMa && rate <= adaREr) {
            if (flagType) {
                if (mis <= misTmp && totalMap >= totalMapTmp) { startPos = r1; misTmp = mis; totalMapTmp = totalMap; }
            } else { startPos = r1; flagType = 1; misTmp = mis; totalMapTmp = totalMap; }
        }
        if (a1 > 0) a1--; else r1++;
    }
    return startPos;
}

/* read_filter.cpp:863-926 sRNA_hasAdapter(): is the 5' adapter's tail (at least adaRCtg bases) present?
 * adapter offsets adptLen-adaRCtg .. 0 at read position 0, then offset 0 at read positions
 * 1..max(0,readLen-adaRCtg); accepts at most 4 mismatches, a run of adaRCtg matches (or a read shorter
 * than 12) and a match fraction of adaRAr relative to the read or to the adapter. */
ORC_EXPORT int orc_srna_has_adapter(const uint8_t* read, int readLen, const uint8_t* adapter, int adptLen,
                                    int adaRCtg, float adaRAr)
{
    if (adptLen == 0) return 0;
    int a1 = adptLen - adaRCtg;
    int readLenSmall = (readLen - adaRCtg < 0) ? 0 : (readLen - adaRCtg);
    for (int r1 = 0; r1 <= readLenSmall;) {
        int len1 = adptLen - a1, len2 = readLen - r1;
        int len = len1 < len2 ? len1 : len2;
        int mis = 0, totalMap = 0, run = 0, max_map = 0;
        for (int c = 0; c < len; c++) {
            if (adapter[a1 + c] == read[r1 + c]) { totalMap++; if (++run > max_map) max_map = run; }
            else { mis++; run = 0; }
        }
        if (mis <= 4 && (max_map >= adaRCtg || readLen < 12) &&
            (1.0 * totalMap / readLen >= adaRAr || 1.0 * totalMap / adptLen >= adaRAr))
            return 1;
        if (a1 > 0) a1--; else r1++;
    }
    return 0;
}

/* read_filter.cpp:596-706 hasContam() (the list variant :507-595 only differs in where segMatchThr comes
 * from: the caller passes it). Three phases like adapter_pos, but: an 'N' of the read is neither a match
 * nor a mismatch (and does not break a run of matches), phases 1 and 3 accept a run of segMatchTemp =
 * 7 + r1/segGrad matches, and the mismatch budget grows with r1/misGrad. Float quotients are assigned to
 * int as on x86 (NaN / inf -> INT_MIN). Read positions outside the read compare as mismatches (the
 * reference reads out of bounds there). adaMis / adaEdge are the mate's own: read 2 is analysed with a
 * parameter copy that carries adaMis2 / adaEdge2 (sequence.cpp:183-188). */
ORC_EXPORT int orc_has_contam(const uint8_t* read, int readLen, const uint8_t* contam, int contamLen, int segMatchThr,
                              int adaMis, int adaEdge)
{
    if (contamLen == 0) return -1;
    float misGrad = (float)((contamLen - adaEdge) / (adaMis + 1));
    float segGrad = (segMatchThr - 7 + 1 == 0) ? 0.0f : (float)((contamLen - adaEdge) / (segMatchThr - 7 + 1));
    int r1, mis, seg, misT, segT;
    for (r1 = 0; r1 < contamLen - adaEdge; ++r1) {                       /* contaminant's tail at the read's head */
        mis = 0; seg = 0;
        misT = float_to_int_x86((float)r1 / misGrad);
        segT = segGrad != 0 ? float_to_int_x86(7 + (float)r1 / segGrad) : 7;
        for (int c = 0; c < r1 + adaEdge; ++c) {
            int rc = c < readLen ? read[c] : -1;
            if (rc >= 0 && contam[contamLen - r1 - adaEdge + c] == rc) { if (++seg >= segT) return 0; }
            else if (rc != 'N') { mis++; seg = 0; if (mis > misT) break; }
        }
        if (mis <= misT) return 0;
    }
    for (r1 = 0; r1 <= readLen - contamLen; ++r1) {                      /* whole contaminant inside the read */
        seg = 0; mis = 0;
        for (int c = 0; c < contamLen; ++c) {
            if (contam[c] == read[r1 + c]) { if (++seg >= segMatchThr) return r1; }
            else if (read[r1 + c] != 'N') { mis++; seg = 0; if (mis > adaMis) break; }
        }
        if (mis <= adaMis) return r1;
    }
    for (r1 = 0; r1 < contamLen - adaEdge; ++r1) {                       /* contaminant's head at the read's tail */
        mis = 0; seg = 0;
        misT = float_to_int_x86((float)r1 / misGrad);
        segT = float_to_int_x86(7 + (float)r1 / segGrad);               /* no zero test here (:686) */
        int base = readLen - r1 - adaEdge;
        for (int c = 0; c < r1 + adaEdge; ++c) {
            int idx = base + c;
            int rc = (idx >= 0 && idx < readLen) ? read[idx] : -1;
            if (rc >= 0 && contam[c] == rc) { if (++seg >= segT) return base; }
            else if (rc != 'N') { mis++; seg = 0; if (mis > misT) break; }
        }
        if (mis <= misT) return base;
    }
    return -1;
}

/* read_filter.cpp:961-1053 global_contam_pos(): a scoring walk (match +1, mismatch -200) over three
 * placements of the contaminant; total_score / overlap are NOT reset between the start positions of the
 * second and third placement, only before each placement. */
ORC_EXPORT int orc_global_contam_pos(const uint8_t* read, int rl, const uint8_t* ct, int cl, int min_match_len, int mismatch_number)
{
    const int mismatch_score = -200, match_score = 1;
    int total_mismatch_score = mismatch_number * mismatch_score;
    int lower_score = (min_match_len - mismatch_number) + total_mismatch_score;
    int total_score = -1000, overlap = 0;
    for (int i = cl - min_match_len; i >= 0; i--) {                       /* contaminant in front of the read */
        int j_max = cl - i > rl ? rl : cl - i;
        for (int j = 0; j != j_max; j++) {
            if (read[j] == ct[i + j]) {
                if (total_score > total_mismatch_score) { total_score += match_score; overlap++; }
                else { if (j_max - j < min_match_len) break; total_score = match_score; overlap = 1; }
            } else {
                if (total_score > total_mismatch_score) { total_score += mismatch_score; overlap++; }
                else if (j_max - j < min_match_len) break;
            }
            if (total_score >= lower_score && overlap >= min_match_len) return 0;
        }
    }
    total_score = -1000; overlap = 0;
    for (int i = 0; i <= rl - cl; i++) {                                  /* in the middle */
        for (int j = 0; j != cl; j++) {
            if (read[i + j] == ct[j]) {
                if (total_score > total_mismatch_score) { total_score += match_score; overlap++; }
                else { if (cl - j < min_match_len) break; total_score = match_score; overlap = 1; }
            } else {
                if (total_score > total_mismatch_score) { total_score += mismatch_score; overlap++; }
                else if (cl - j < min_match_len) break;
            }
            if (total_score >= lower_score && overlap >= min_match_len) return i + j - overlap + 1;
        }
    }
    total_score = -1000; overlap = 0;
    int i_min = cl > rl ? cl - rl : 0;
    for (int i = i_min; i <= cl - min_match_len; i++) {                   /* at the tail */
        for (int j = 0; j != cl - i; j++) {
            if (read[rl - (cl - i) + j] == ct[j]) {
                if (total_score > total_mismatch_score) { total_score += match_score; overlap++; }
                else { total_score = match_score; overlap = 1; if (cl - i - j < min_match_len) break; }
            } else {
                if (total_score > total_mismatch_score) { total_score += mismatch_score; overlap++; }
                else if (cl - i - j < min_match_len) break;
            }
            if (total_score >= lower_score && overlap >= min_match_len) return rl - cl + i + j - overlap + 1;
        }
    }
    return -1;
}
/* read_filter.cpp:1055-1075 reversecomplementary() of a contaminant; returns 0 on an unrecognized base */
static int orc_revcomp(const char* a, int n, uint8_t* out)
{
    for (int k = 0; k < n; k++) {
        int ch = a[n - 1 - k];
        if (ch >= 'a' && ch <= 'z') ch -= 32;
        switch (ch) {
            case 'A': out[k] = 'T'; break; case 'T': out[k] = 'A'; break;
            case 'G': out[k] = 'C'; break; case 'C': out[k] = 'G'; break;
            case 'N': out[k] = 'N'; break;
            default: return 0;
        }
    }
    return 1;
}
/* :207-249 of stat_read restricted to what the discard uses: include_global_contam = some sequence, forward
 * or reverse complemented, is found (hasGlobalContams :927-960; its early break cannot change that) */
static int orc_has_global_contam(const snk_params* p, const uint8_t* seq, int len)
{
    for (int i = 0; i < p->n_gcontams; i++) {
        const int cl = p->gcontam_len[i];
        uint8_t rev[SNK_MAX_ADAPTER_LEN];
        if (orc_global_contam_pos(seq, len, (const uint8_t*)p->gcontam[i], cl, p->gcontam_min_match[i], p->gcontam_mismatch[i]) >= 0) return 1;
        if (orc_revcomp(p->gcontam[i], cl, rev) &&
            orc_global_contam_pos(seq, len, rev, cl, p->gcontam_min_match[i], p->gcontam_mismatch[i]) >= 0) return 1;
    }
    return 0;
}

/* stat_read's id parse (read_filter.cpp:86-148) + check_tile_or_fov (:14-79): SNK_PRE_TILE / SNK_PRE_FOV
 * bits for one record id (idlen bytes, no terminator needed). The tile is the (up to) 4 digits behind the
 * 2nd ':' (seqType 0) or the 4th ':' (seqType 1); the fov the 8 bytes from the first 'C' that has an 'R'
 * four bytes later (and at least 9 bytes to the end of the id). A read is selected when that string equals
 * an entry of the removal list (ranges with '-' never match anything in the reference). */
ORC_EXPORT uint32_t orc_id_flags(const snk_params* p, const uint8_t* id, int idlen)
{
    uint32_t flags = 0;
    if (p->n_tile > 0) {
        int i = 0, num = 0;
        const int want = p->seq_type1 ? 4 : 2;
        for (; i < idlen; i++) {
            if (id[i] == ':') num++;
            if (num >= want) break;
        }
        char tile[5];
        int tn = 0;
        for (int j = 0; j != 4; j++) {
            int k = i + j + 1;
            int ch = k < idlen ? id[k] : 0;                      /* std::string: id[size()] is '\0', beyond is UB */
            if (ch >= '0' && ch <= '9') tile[tn++] = (char)ch;
        }
        tile[tn] = 0;
        for (int e = 0; e < p->n_tile; e++) {
            int el = (int)strnlen(p->tile[e], SNK_ID_FILTER_LEN);
            if (el == tn && memcmp(p->tile[e], tile, (size_t)tn) == 0) { flags |= SNK_PRE_TILE; break; }
        }
    }
    if (p->n_fov > 0) {
        int i = 0;
        for (i = 0; i < idlen; i++)
            if (id[i] == 'C' && i + 8 < idlen && id[i + 4] == 'R') break;
        int fn = idlen - i; if (fn > 8) fn = 8; if (fn < 0) fn = 0;   /* substr(i, 8) */
        for (int e = 0; e < p->n_fov; e++) {
            int el = (int)strnlen(p->fov[e], SNK_ID_FILTER_LEN);
            if (el == fn && memcmp(p->fov[e], id + i, (size_t)fn) == 0) { flags |= SNK_PRE_FOV; break; }
        }
    }
    return flags;
}

/* everything stat_read / fastq_trim leave behind for one mate */
typedef struct {
    int len;
    int a, c, g, t, n;
    int contig;
    float n_ratio, a_ratio, lowq_ratio, mean_q;
    int has_adapter;
    int include_3_adapter;      /* filtersRNA: sRNA_findAdapter() of the raw read */
    int include_contam;         /* some contaminant of the mate's list was found */
    int include_global_contam;  /* some global contaminant (either strand) was found */
    /* C_fastq cut bookkeeping (sequence.h:69), -1 as set by C_fastq_init (peprocess.cpp:1674-1689) */
    int head_hdcut, head_lqcut, tail_hdcut, tail_lqcut, adacut_pos;
    int head_cut, clean_len;    /* result of fastq_trim */
    int bad_base;               /* unrecognized base seen */
} orc_read;

static int trimming_enabled(const snk_params* p)
{
    /* read_filter.cpp:343-355 */
    return p->has_hard_trim || p->has_trim_bad_head || p->has_trim_bad_tail || p->index_remove ||
           p->ada_trim || p->contam_trim || p->polyG_tail != -1;
}
static int cutback_enabled(const snk_params* p)
{
    /* peprocess.cpp:1441, seprocess.cpp:881 */
    return p->ada_trim || p->contam_trim || p->has_hard_trim || p->has_trim_bad_head || p->has_trim_bad_tail;
}

/* read_filter.cpp:80-313 stat_read (adapter + counters) followed by :338-471 fastq_trim */
static void orc_stat_and_trim(const snk_params* p, int mate, const uint8_t* seq, const uint8_t* qual,
                              int len, orc_read* r)
{
    memset(r, 0, sizeof(*r));
    r->len = len;
    r->head_hdcut = r->head_lqcut = r->tail_hdcut = r->tail_lqcut = r->adacut_pos = -1;
    /* :175-188 first adapter in the list that hits wins */
    int ada_pos = -1;
    if (p->srna) {
        /* :170-174: 3' adapter = adapter2_seq, 5' adapter = adapter1_seq; adacut_pos stays -1 */
        r->include_3_adapter = orc_srna_find_adapter(seq, len, (const uint8_t*)p->adapter[1][0], p->n_adapters[1] ? p->adapter_len[1][0] : 0,
                                                     p->ada_rma, p->ada_rmm, p->ada_rer);
        r->has_adapter = orc_srna_has_adapter(seq, len, (const uint8_t*)p->adapter[0][0], p->n_adapters[0] ? p->adapter_len[0][0] : 0,
                                              p->ada_rctg, p->ada_rar);
    } else {
        for (int i = 0; i < p->n_adapters[mate]; i++) {
            ada_pos = orc_adapter_pos(seq, len, (const uint8_t*)p->adapter[mate][i], p->adapter_len[mate][i],
                                      p->ada_mis[mate], p->ada_mr[mate], p->ada_edge[mate]);
            if (ada_pos >= 0) break;
        }
        if (ada_pos >= 0) { r->has_adapter = 1; r->adacut_pos = len - ada_pos; }
        /* :188-206 hasContam / hasContams: any hit sets include_contam (the list's early break cannot change that) */
        for (int i = 0; i < p->n_contams[mate]; i++)
            if (orc_has_contam(seq, len, (const uint8_t*)p->contam[mate][i], p->contam_len[mate][i], p->contam_seg_thr[mate][i],
                               p->ada_mis[mate], p->ada_edge[mate]) >= 0) { r->include_contam = 1; break; }
        r->include_global_contam = orc_has_global_contam(p, seq, len);
    }
    /* :255-287 base loop */
    int last_char = 'Q', contig = 0, max_contig = 1;
    for (int i = 0; i < len; i++) {
        int ch = seq[i];
        if (p->polyX_num != -1) {
            if (ch == last_char) { contig++; if (max_contig < contig) max_contig = contig; }
            else contig = 1;
        }
        last_char = ch;
        switch (ch) {
            case 'a': case 'A': r->a++; break;
            case 'c': case 'C': r->c++; break;
            case 'g': case 'G': r->g++; break;
            case 't': case 'T': r->t++; break;
            case 'n': case 'N': r->n++; break;
            default: r->bad_base = 1; break;       /* :282-285 exit(1) */
        }
    }
    r->contig = max_contig;
    r->a_ratio = (float)r->a / (float)len;          /* :290 */
    r->n_ratio = (float)r->n / (float)len;          /* :294 */
    /* :296-311 quality loop */
    int total = 0, low = 0;
    for (int i = 0; i < len; i++) {
        int q = (int)qual[i] - p->quality_phred;
        total += q;
        if (q <= p->low_qual) low++;
    }
    r->lowq_ratio = (float)low / (float)len;
    r->mean_q = (float)total / (float)len;

    /* ---- fastq_trim, read_filter.cpp:338-471 ---- */
    r->head_cut = 0; r->clean_len = len;
    if (!trimming_enabled(p)) return;               /* :354-355 */
    int head_cut = 0, tail_cut = 0;
    if (p->has_hard_trim) {                         /* :384-389 */
        r->head_hdcut = p->hard_head[mate];
        r->tail_hdcut = p->hard_tail[mate];
        head_cut = r->head_hdcut; tail_cut = r->tail_hdcut;
    }
    if (p->has_trim_bad_head || p->has_trim_bad_tail) {   /* :390-429 */
        int hthr = p->has_trim_bad_head ? p->bad_head_thr : 0, hmax = p->has_trim_bad_head ? p->bad_head_max : 0;
        int tthr = p->has_trim_bad_tail ? p->bad_tail_thr : 0, tmax = p->has_trim_bad_tail ? p->bad_tail_max : 0;
        int hix = 0, tix = 0;
        for (int ix = 0; ix < hmax && ix < len; ix++) {   /* bounded by len: reference reads past the end (UB) */
            if ((int)qual[ix] - p->quality_phred < hthr) hix++; else break;
        }
        for (int ix = 0; ix < tmax && ix < len; ix++) {
            if ((int)qual[len - ix - 1] - p->quality_phred < tthr) tix++; else break;
        }
        r->head_lqcut = hix; r->tail_lqcut = tix;
        if (hix > head_cut) head_cut = hix;
        if (tix > tail_cut) tail_cut = tix;
    }
    int cur = len;                                  /* read.sequence.size() as the function goes on */
    if (p->ada_trim) {                              /* :430-442 */
        if (p->srna) {
            /* :432-438 the read is cut at the 3' adapter BEFORE the head/tail cuts are applied (the same
               search as in stat_read: the sequence has not been modified yet) */
            int pos3 = r->include_3_adapter;
            if (pos3 > 2 && pos3 < cur) cur = pos3;
        }
        if (r->adacut_pos > 0 && r->adacut_pos > tail_cut) tail_cut = r->adacut_pos;
    }
    if (p->polyG_tail != -1) {                      /* :454-461, polyG_number :472-482 */
        int ng = 0;
        for (int i = cur - 1; i >= 0; i--) { if (seq[i] == 'G' || seq[i] == 'g') ng++; else break; }
        if ((float)ng >= p->polyG_tail) { if (ng > tail_cut) tail_cut = ng; }
    }
    if (head_cut + tail_cut > cur) { r->head_cut = 0; r->clean_len = 0; }   /* :462-464 */
    else { r->head_cut = head_cut; r->clean_len = cur - head_cut - tail_cut; }
}

static void ts_bump(uint64_t* ts, int arr, long idx)
{
    long flat = (long)arr * SNK_MAX_READ_LEN + idx;     /* negative idx spills into the previous array */
    if (flat >= 0 && flat < SNK_TS_WORDS) ts[flat]++;
}

/* peprocess.cpp:1105-1204 (fq1), :1323-1421 (fq2), seprocess.cpp:645-740: one record into one table.
 * which: 0 = PE fq1 (index base raw_length), 1 = PE fq2 (index base sequence.size()), 2 = SE */
static void orc_stat_record(const snk_params* p, uint64_t* file, int which, const uint8_t* seq,
                            const uint8_t* qual, int slen, int raw_length,
                            int head_hdcut, int head_lqcut, int tail_hdcut, int tail_lqcut, int adacut_pos,
                            uint64_t key_index, uint32_t* err)
{
    uint64_t* gs = file + SNK_FILE_GS_OFF;
    uint64_t* bs = file + SNK_FILE_BS_OFF;
    uint64_t* qs = file + SNK_FILE_QS_OFF;
    uint64_t* ts = file + SNK_FILE_TS_OFF;
    if (head_hdcut > 0 || head_lqcut > 0) {
        if (head_hdcut >= head_lqcut) ts_bump(ts, SNK_TS_HT, head_hdcut);
        else ts_bump(ts, SNK_TS_HLQ, head_lqcut);
    }
    int ada_cond = (which == 2) ? (adacut_pos >= 0) : (adacut_pos > 0);   /* seprocess.cpp:658 vs peprocess.cpp:1118 */
    if (tail_hdcut > 0 || tail_lqcut > 0 || ada_cond) {
        long base = (which == 1) ? slen : raw_length;
        if (tail_hdcut >= tail_lqcut) {
            if (tail_hdcut >= adacut_pos) ts_bump(ts, SNK_TS_TT, base - tail_hdcut + 1);
            else ts_bump(ts, SNK_TS_TA, base - adacut_pos + 1);
        } else {
            if (tail_lqcut >= adacut_pos) ts_bump(ts, SNK_TS_TLQ, base - tail_lqcut + 1);
            else ts_bump(ts, SNK_TS_TA, base - adacut_pos + 1);
        }
    }
    for (int i = 0; i < slen; i++) {
        switch (seq[i]) {
            case 'a': case 'A': bs[i * 5 + 0]++; gs[SNK_GS_A]++; break;
            case 'c': case 'C': bs[i * 5 + 1]++; gs[SNK_GS_C]++; break;
            case 'g': case 'G': bs[i * 5 + 2]++; gs[SNK_GS_G]++; break;
            case 't': case 'T': bs[i * 5 + 3]++; gs[SNK_GS_T]++; break;
            case 'n': case 'N': bs[i * 5 + 4]++; gs[SNK_GS_N]++; break;
            default: *err |= 1u; break;
        }
        int q = (int)qual[i] - p->quality_phred;    /* clean uses rebased char - outputQualityPhred == same q */
        if (q < 0 || q >= SNK_QBINS) *err |= 2u;
        else qs[(size_t)i * SNK_QBINS + q]++;
        if (q >= 20) gs[SNK_GS_Q20]++;
        if (q >= 30) gs[SNK_GS_Q30]++;
    }
    gs[SNK_GS_BASES] += (uint64_t)slen;
    gs[SNK_GS_READS] += 1;
    uint64_t key = ((key_index + 1) << 16) | (uint64_t)slen;
    if (key > gs[SNK_GS_LAST_KEY]) gs[SNK_GS_LAST_KEY] = key;
}

static void fs_dis(uint64_t* fs, int base, int a, int b)
{
    /* pe_dis + switch, sequence.cpp:392-399 and e.g. :292-302 */
    if (a) fs[base + 1]++;
    if (b) fs[base + 2]++;
    if (a && b) fs[base + 3]++;
    fs[base]++;
}

/* sequence.cpp:198-387 pe_discard (contam/tile/fov/dup/overlap branches are out of scope) */
static int orc_pe_discard(const snk_params* p, const orc_read* r1, const orc_read* r2, uint32_t pre, uint64_t* fs,
                          int* mask, uint32_t* err)
{
    int a, b;
    /* :213-230 tile / fov of fastq1 only, plain counters */
    if (pre & SNK_PRE_TILE) { fs[SNK_FS_TILE]++; *mask = 0; return SNK_DROP_TILE; }
    if (pre & SNK_PRE_FOV) { fs[SNK_FS_FOV]++; *mask = 0; return SNK_DROP_FOV; }
#define DIS(cat, base) do { if (a || b) { fs_dis(fs, base, a, b); *mask = (a ? 1 : 0) | (b ? 2 : 0); return cat; } } while (0)
    if (p->min_read_length != -1) {
        /* size_t < int comparison: negative thresholds other than -1 convert to huge unsigned */
        a = (uint64_t)r1->clean_len < (uint64_t)(int64_t)p->min_read_length;
        b = (uint64_t)r2->clean_len < (uint64_t)(int64_t)p->min_read_length;
        DIS(SNK_DROP_SHORT, SNK_FS_SHORT);
    } else if (r1->clean_len == 0 || r2->clean_len == 0) { *mask = 0; return SNK_DROP_EMPTY; }
    if (p->max_read_length != -1) {
        a = (uint64_t)r1->clean_len > (uint64_t)(int64_t)p->max_read_length;
        b = (uint64_t)r2->clean_len > (uint64_t)(int64_t)p->max_read_length;
        DIS(SNK_DROP_LONG, SNK_FS_LONG);
    }
    if (p->contam_discard) { a = r1->include_global_contam; b = r2->include_global_contam; DIS(SNK_DROP_GCONTAM, SNK_FS_GCONTAM); }   /* :262-273 */
    if (p->contam_discard) { a = r1->include_contam; b = r2->include_contam; DIS(SNK_DROP_CONTAM, SNK_FS_CONTAM); }   /* :274-288 */
    if (p->n_ratio != -1) { a = r1->n_ratio >= p->n_ratio; b = r2->n_ratio >= p->n_ratio; DIS(SNK_DROP_N, SNK_FS_N); }
    if (p->highA_ratio != -1) { a = r1->a_ratio >= p->highA_ratio; b = r2->a_ratio >= p->highA_ratio; DIS(SNK_DROP_HIGHA, SNK_FS_HIGHA); }
    if (p->polyX_num != -1) { a = r1->contig >= p->polyX_num; b = r2->contig >= p->polyX_num; DIS(SNK_DROP_POLYX, SNK_FS_POLYX); }
    if (p->low_qual_ratio != -1) {
        a = r1->lowq_ratio >= p->low_qual_ratio; b = r2->lowq_ratio >= p->low_qual_ratio;
        if ((a || b) && (r1->lowq_ratio > 1 || r2->lowq_ratio > 1)) *err |= 4u;   /* :335-338 */
        DIS(SNK_DROP_LOWQ, SNK_FS_LOWQ);
    }
    if (p->mean_quality != -1) { a = r1->mean_q < (float)p->mean_quality; b = r2->mean_q < (float)p->mean_quality; DIS(SNK_DROP_MEANQ, SNK_FS_MEANQ); }
    if (!p->ada_trim) { a = r1->has_adapter; b = r2->has_adapter; DIS(SNK_DROP_ADAPTER, SNK_FS_ADAPTER); }
#undef DIS
    *mask = 0;
    return SNK_KEEP;
}

/* sequence.cpp:76-178 se_discard */
static int orc_se_discard(const snk_params* p, const orc_read* r, uint32_t pre, uint64_t* fs)
{
    if (pre & SNK_PRE_TILE) { fs[SNK_FS_TILE]++; return SNK_DROP_TILE; }      /* :84-98 */
    if (pre & SNK_PRE_FOV) { fs[SNK_FS_FOV]++; return SNK_DROP_FOV; }
    if (p->min_read_length != -1 && (uint64_t)r->clean_len < (uint64_t)(int64_t)p->min_read_length) { fs[SNK_FS_SHORT]++; return SNK_DROP_SHORT; }
    if (p->max_read_length != -1 && (uint64_t)r->clean_len > (uint64_t)(int64_t)p->max_read_length) { fs[SNK_FS_LONG]++; return SNK_DROP_LONG; }
    if (p->contam_discard && r->include_contam) { fs[SNK_FS_CONTAM]++; return SNK_DROP_CONTAM; }                       /* :116-128 */
    if (p->contam_discard && r->include_global_contam) { fs[SNK_FS_GCONTAM]++; return SNK_DROP_GCONTAM; }
    if (p->n_ratio != -1 && r->n_ratio >= p->n_ratio) { fs[SNK_FS_N]++; return SNK_DROP_N; }
    if (p->highA_ratio != -1 && r->a_ratio >= p->highA_ratio) { fs[SNK_FS_HIGHA]++; return SNK_DROP_HIGHA; }
    if (p->polyX_num != -1 && r->contig >= p->polyX_num) { fs[SNK_FS_POLYX]++; return SNK_DROP_POLYX; }
    if (p->low_qual_ratio != -1 && r->lowq_ratio >= p->low_qual_ratio) { fs[SNK_FS_LOWQ]++; return SNK_DROP_LOWQ; }
    if (p->mean_quality != -1 && r->mean_q < (float)p->mean_quality) { fs[SNK_FS_MEANQ]++; return SNK_DROP_MEANQ; }
    if (r->has_adapter && !p->ada_trim) { fs[SNK_FS_ADAPTER]++; return SNK_DROP_ADAPTER; }
    return SNK_KEEP;
}

/* sequence.cpp:19-75 sRNA_discard */
static int orc_srna_discard(const snk_params* p, const orc_read* r, uint64_t* fs, uint32_t* err)
{
    if (p->max_read_length != -1 && (uint64_t)r->clean_len > (uint64_t)(int64_t)p->max_read_length) { fs[SNK_FS_LONG]++; return SNK_DROP_LONG; }
    if (p->low_qual_ratio != -1 && r->lowq_ratio >= p->low_qual_ratio) { fs[SNK_FS_LOWQ]++; return SNK_DROP_LOWQ; }
    if (r->include_3_adapter == -1) { fs[SNK_FS_NO3ADAPTER]++; return SNK_DROP_NO3ADAPTER; }
    if (r->include_3_adapter <= 2) { fs[SNK_FS_INSERTNULL]++; return SNK_DROP_INSERTNULL; }
    if (r->has_adapter) { fs[SNK_FS_ADAPTER]++; return SNK_DROP_ADAPTER; }
    if (p->highA_ratio != -1 && r->a_ratio >= p->highA_ratio) { fs[SNK_FS_HIGHA]++; return SNK_DROP_HIGHA; }
    if (p->polyX_num != -1 && r->contig >= p->polyX_num) { fs[SNK_FS_POLYX]++; return SNK_DROP_POLYX; }
    if ((uint64_t)r->clean_len < (uint64_t)(int64_t)p->min_read_length) { fs[SNK_FS_SHORT]++; return SNK_DROP_SHORT; }   /* :68 no -1 test */
    (void)err;
    return SNK_KEEP;
}

static void fill_result(snk_read_result* o, const orc_read* r, int cat, int mask)
{
    o->head_cut = (uint16_t)r->head_cut;
    o->clean_len = (uint16_t)r->clean_len;
    o->category = (uint8_t)cat;
    o->mate_mask = (uint8_t)mask;
    o->adacut_pos = (int16_t)r->adacut_pos;
}

/* filter_pe_fqs (peprocess.cpp:1424-1484) + stat_pe_fqs raw (:1923) + stat_pe_fqs clean (:1961).
 * stats: n_slots blocks of SNK_SLOT_WORDS uint64, accumulated into. err: sticky error bits. */
ORC_EXPORT int orc_filter_pe(const snk_params* p, const snk_batch* b1, const snk_batch* b2,
                             snk_read_result* out1, snk_read_result* out2, uint64_t* stats,
                             uint64_t first_index, uint32_t* err)
{
    if (b1->n != b2->n) return 1;
    int cutback = cutback_enabled(p);
    for (uint32_t i = 0; i < b1->n; i++) {
        uint64_t gi = first_index + i;
        int slot = (int)((gi / (uint64_t)p->slot_block) % (uint64_t)p->n_slots);
        uint64_t* S = stats + (size_t)slot * SNK_SLOT_WORDS;
        const uint8_t* s1 = b1->seq + (size_t)i * b1->stride; const uint8_t* q1 = b1->qual + (size_t)i * b1->stride;
        const uint8_t* s2 = b2->seq + (size_t)i * b2->stride; const uint8_t* q2 = b2->qual + (size_t)i * b2->stride;
        orc_read r1, r2;
        orc_stat_and_trim(p, 0, s1, q1, b1->len[i] & SNK_LEN_MASK, &r1);
        orc_stat_and_trim(p, 1, s2, q2, b2->len[i] & SNK_LEN_MASK, &r2);
        if (r1.bad_base || r2.bad_base) *err |= 1u;
        int mask = 0;
        int cat = orc_pe_discard(p, &r1, &r2, b1->len[i] & ~SNK_LEN_MASK, S, &mask, err);
        fill_result(&out1[i], &r1, cat, mask);
        fill_result(&out2[i], &r2, cat, mask);
        /* raw tables: raw records keep raw_length==0 and -1 cuts unless copied back (peprocess.cpp:1441-1459) */
        orc_stat_record(p, S + SNK_SLOT_FILE_OFF(SNK_RAW1), 0, s1, q1, r1.len, 0,
                        cutback ? r1.head_hdcut : -1, cutback ? r1.head_lqcut : -1, cutback ? r1.tail_hdcut : -1,
                        cutback ? r1.tail_lqcut : -1, cutback ? r1.adacut_pos : -1, gi, err);
        orc_stat_record(p, S + SNK_SLOT_FILE_OFF(SNK_RAW2), 1, s2, q2, r2.len, 0,
                        cutback ? r2.head_hdcut : -1, cutback ? r2.head_lqcut : -1, cutback ? r2.tail_hdcut : -1,
                        cutback ? r2.tail_lqcut : -1, cutback ? r2.adacut_pos : -1, gi, err);
        if (cat == SNK_KEEP) {
            /* clean records are the filter's trimmed copies: real raw_length, all bookkeeping ints */
            orc_stat_record(p, S + SNK_SLOT_FILE_OFF(SNK_CLEAN1), 0, s1 + r1.head_cut, q1 + r1.head_cut, r1.clean_len, r1.len,
                            r1.head_hdcut, r1.head_lqcut, r1.tail_hdcut, r1.tail_lqcut, r1.adacut_pos, gi, err);
            orc_stat_record(p, S + SNK_SLOT_FILE_OFF(SNK_CLEAN2), 1, s2 + r2.head_cut, q2 + r2.head_cut, r2.clean_len, r2.len,
                            r2.head_hdcut, r2.head_lqcut, r2.tail_hdcut, r2.tail_lqcut, r2.adacut_pos, gi, err);
        }
    }
    return 0;
}

/* filter_se_fqs (seprocess.cpp:871-917) + stat_se_fqs raw (:1967) + clean (:2007) */
ORC_EXPORT int orc_filter_se(const snk_params* p, const snk_batch* b1, snk_read_result* out1,
                             uint64_t* stats, uint64_t first_index, uint32_t* err)
{
    int cutback = cutback_enabled(p);
    for (uint32_t i = 0; i < b1->n; i++) {
        uint64_t gi = first_index + i;
        int slot = (int)((gi / (uint64_t)p->slot_block) % (uint64_t)p->n_slots);
        uint64_t* S = stats + (size_t)slot * SNK_SLOT_WORDS;
        const uint8_t* s1 = b1->seq + (size_t)i * b1->stride; const uint8_t* q1 = b1->qual + (size_t)i * b1->stride;
        orc_read r1;
        orc_stat_and_trim(p, 0, s1, q1, b1->len[i] & SNK_LEN_MASK, &r1);
        if (r1.bad_base) *err |= 1u;
        int cat = p->srna ? orc_srna_discard(p, &r1, S, err) : orc_se_discard(p, &r1, b1->len[i] & ~SNK_LEN_MASK, S);
        fill_result(&out1[i], &r1, cat, cat ? 1 : 0);
        orc_stat_record(p, S + SNK_SLOT_FILE_OFF(SNK_RAW1), 2, s1, q1, r1.len, 0,
                        cutback ? r1.head_hdcut : -1, cutback ? r1.head_lqcut : -1, cutback ? r1.tail_hdcut : -1,
                        cutback ? r1.tail_lqcut : -1, cutback ? r1.adacut_pos : -1, gi, err);
        if (cat == SNK_KEEP)
            orc_stat_record(p, S + SNK_SLOT_FILE_OFF(SNK_CLEAN1), 2, s1 + r1.head_cut, q1 + r1.head_cut, r1.clean_len, r1.len,
                            r1.head_hdcut, r1.head_lqcut, r1.tail_hdcut, r1.tail_lqcut, r1.adacut_pos, gi, err);
    }
    return 0;
}

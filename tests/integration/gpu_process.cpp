// gpu_process.cpp — see gpu_process.h. Own code written against two public interfaces: the reference's
// peProcess / seProcess class declarations and include/snk_engine.h.
#include "gpu_process.h"
#include "cli_params.h"          // snk::HostParams / to_engine_params: the same gp -> snk_params mapping the drop-in CLI uses
#include "text_core.cuh"         // snkcore::id_transform (host-compilable): index removal of fastq_trim (read_filter.cpp:357-382)
#include <cstring>
#include <mutex>
#include <vector>

namespace {

snk::HostParams host_params_of(const C_global_parameter& gp, bool pe)
{
    snk::HostParams hp;
    hp.module_name = gp.module_name;
    hp.is_pe = pe;
    hp.seq_type = gp.seq_type;
    hp.ada_trim = gp.adapter_discard_or_trim == "trim";
    hp.ada1s = gp.ada1s; hp.ada2s = gp.ada2s;
    hp.quality_phred = gp.qualityPhred; hp.out_quality_phred = gp.outputQualityPhred;
    hp.low_qual = gp.lowQual; hp.low_qual_ratio = gp.lowQualityBaseRatio; hp.mean_quality = gp.meanQuality;
    hp.trim_bad_head = gp.trimBadHead; hp.trim_bad_tail = gp.trimBadTail; hp.trim = gp.trim;
    hp.max_base_quality = gp.maxBaseQuality;
    hp.n_ratio = gp.n_ratio; hp.highA_ratio = gp.highA_ratio; hp.polyG_tail = gp.polyG_tail; hp.polyX_num = gp.polyX_num;
    hp.index_remove = gp.index_remove;
    hp.threads = hp.threads_requested = gp.threads_num; hp.patch_size = gp.patchSize;
    hp.max_read_length = gp.max_read_length; hp.min_read_length = gp.min_read_length;
    hp.ada_mis = gp.adaMis; hp.ada_mis2 = gp.adaMis2; hp.ada_edge = gp.adaEdge; hp.ada_edge2 = gp.adaEdge2;
    hp.ada_mr = gp.adaMR; hp.ada_mr2 = gp.adaMR2;
    hp.contam1 = gp.contam1_seq; hp.contam2 = gp.contam2_seq; hp.ct_match_r = gp.ctMatchR;
    hp.contam_trim = gp.contam_discard_or_trim == "trim";
    hp.global_contams = gp.global_contams; hp.g_mrs = gp.g_mrs; hp.g_mms = gp.g_mms;
    hp.tile = gp.tile; hp.fov = gp.fov;
    hp.srna = gp.module_name == "filtersRNA";
    hp.ada_rctg = gp.adaRCtg; hp.ada_rma = gp.adaRMa; hp.ada_rmm = gp.adaRMm; hp.ada_rar = gp.adaRAr; hp.ada_rer = gp.adaREr;
    if (hp.srna) { hp.ada1s = {gp.adapter1_seq}; hp.ada2s = {gp.adapter2_seq}; }      // 5' / 3' adapter of filtersRNA
    return hp;
}

[[noreturn]] void engine_die() { cerr << "Error:" << snk_last_error() << endl; exit(1); }

// C_fastq records -> one mate of a fixed-stride SoA batch
struct Packed {
    std::vector<uint8_t> seq, qual;
    std::vector<uint16_t> len;
    snk_batch b;
    void pack(const vector<C_fastq>& v)
    {
        size_t maxlen = 1;
        for (const C_fastq& f : v) maxlen = std::max(maxlen, f.sequence.size());
        const uint32_t stride = (uint32_t)((maxlen + 15) / 16 * 16);
        seq.assign(v.size() * (size_t)stride, 0); qual.assign(v.size() * (size_t)stride, 0); len.resize(v.size());
        for (size_t i = 0; i < v.size(); i++) {
            if (v[i].sequence.size() != v[i].qual_seq.size()) { cerr << "Error:sequence and quality have different lengths" << endl; exit(1); }
            memcpy(&seq[i * stride], v[i].sequence.data(), v[i].sequence.size());
            memcpy(&qual[i * stride], v[i].qual_seq.data(), v[i].qual_seq.size());
            len[i] = (uint16_t)v[i].sequence.size();
        }
        b.seq = seq.data(); b.qual = qual.data(); b.len = len.data(); b.n = (uint32_t)v.size(); b.stride = stride;
    }
};

// the record as fastq_trim leaves it: cut sequence / qualities, index removed from the id
C_fastq trimmed(const C_fastq& in, const snk_read_result& r, int id_mode)
{
    C_fastq o = in;
    o.sequence = in.sequence.substr(r.head_cut, r.clean_len);
    o.qual_seq = in.qual_seq.substr(r.head_cut, r.clean_len);
    if (id_mode) {
        std::vector<uint8_t> buf(in.seq_id.size() + 1);
        const uint32_t n = snkcore::id_transform((const uint8_t*)in.seq_id.data(), (uint32_t)in.seq_id.size(), id_mode, buf.data());
        o.seq_id.assign((const char*)buf.data(), n);
    }
    return o;
}

// global index of the first record of worker `index`'s next batch: the reference deals blocks of slot_block records
// round robin to its T workers (peprocess.cpp:81, 2063, 2092), each worker walks its blocks in order
struct SlotCursor {
    std::vector<uint64_t> done;
    uint64_t block = 1;
    int T = 1;
    uint64_t first_index(int index, size_t n)
    {
        const uint64_t d = done[index];
        done[index] += n;
        return ((d / block) * (uint64_t)T + (uint64_t)index) * block + d % block;
    }
};

void check_engine_errors(snk_engine* e)
{
    uint32_t flags = 0; uint64_t bad = 0;
    if (snk_engine_error_flags(e, &flags, &bad)) engine_die();
    if (flags & 1) { cerr << "Error:unrecognized sequence, read number " << bad + 1 << endl; exit(1); }
    if (flags & 2) { cerr << "Error:base quality is out of range,please check the quality system parameter or fastq file" << endl; exit(1); }
    if (flags & 4) { cerr << "Error:low quality base ratio stat error" << endl; exit(1); }
    if (flags & 8) { cerr << "Error:read longer than its batch row" << endl; exit(1); }
}

// one file block of a slot's table -> the reference's per-thread accumulator
void fill_file_stat(const uint64_t* F, C_fastq_file_stat& st, int max_q)
{
    const uint64_t* gs = F + SNK_FILE_GS_OFF;
    st.gs.reads_number = gs[SNK_GS_READS]; st.gs.base_number = gs[SNK_GS_BASES];
    st.gs.a_number = gs[SNK_GS_A]; st.gs.c_number = gs[SNK_GS_C]; st.gs.g_number = gs[SNK_GS_G];
    st.gs.t_number = gs[SNK_GS_T]; st.gs.n_number = gs[SNK_GS_N];
    st.gs.q20_num = gs[SNK_GS_Q20]; st.gs.q30_num = gs[SNK_GS_Q30];
    st.gs.read_length = gs[SNK_GS_LAST_KEY] & 0xFFFFu;           // length of the last record the worker saw (peprocess.cpp:1202)
    memcpy(st.bs.position_acgt_content, F + SNK_FILE_BS_OFF, sizeof(uint64_t) * SNK_BS_WORDS);
    for (int pos = 0; pos < READ_MAX_LEN; pos++)
        for (int q = 0; q < max_q && q < SNK_QBINS; q++) st.qs.position_qual[pos][q] = F[SNK_FILE_QS_OFF + (size_t)pos * SNK_QBINS + q];
    // hlq, ht, ta, tlq, tt are consecutive arrays (global_variable.h:118-124), the engine keeps the same layout
    static_assert(sizeof(C_reads_trim_stat) == sizeof(uint64_t) * SNK_TS_WORDS, "C_reads_trim_stat layout");
    memcpy((void*)&st.ts, F + SNK_FILE_TS_OFF, sizeof(uint64_t) * SNK_TS_WORDS);
}

void fill_filter_stat(const uint64_t* S, C_filter_stat& fs)
{
#define SNK_FOUR(field, base) fs.field = S[base]; fs.field##1 = S[base + 1]; fs.field##2 = S[base + 2]; fs.field##_overlap = S[base + 3];
    SNK_FOUR(include_adapter_seq_num, SNK_FS_ADAPTER)
    SNK_FOUR(n_ratio_num, SNK_FS_N)
    SNK_FOUR(highA_num, SNK_FS_HIGHA)
    SNK_FOUR(polyX_num, SNK_FS_POLYX)
    SNK_FOUR(low_qual_base_ratio_num, SNK_FS_LOWQ)
    SNK_FOUR(mean_quality_num, SNK_FS_MEANQ)
    SNK_FOUR(short_len_num, SNK_FS_SHORT)
    SNK_FOUR(long_len_num, SNK_FS_LONG)
    SNK_FOUR(include_contam_seq_num, SNK_FS_CONTAM)
    SNK_FOUR(include_global_contam_seq_num, SNK_FS_GCONTAM)
#undef SNK_FOUR
    fs.no_3_adapter_num = S[SNK_FS_NO3ADAPTER]; fs.int_insertNull_num = S[SNK_FS_INSERTNULL];
    fs.tile_num = S[SNK_FS_TILE]; fs.fov_num = S[SNK_FS_FOV];
}

} // namespace

// ================================================================================================ PE
struct gpuPeProcess::Impl {
    snk_engine* eng = nullptr;
    snk_params params;
    SlotCursor cursor;
    int id_mode = 0;
    bool checked = false;
    std::mutex mu;
};

gpuPeProcess::gpuPeProcess(C_global_parameter m_gp) : peProcess(m_gp), d_(new Impl())
{
    const snk::HostParams hp = host_params_of(gp, true);
    snk::to_engine_params(hp, d_->params);
    if (snk_params_check(&d_->params)) engine_die();
    if (snk_engine_create(&d_->params, 0, &d_->eng)) engine_die();
    d_->cursor.T = gp.threads_num; d_->cursor.block = (uint64_t)d_->params.slot_block; d_->cursor.done.assign(gp.threads_num, 0);
    d_->id_mode = gp.index_remove ? (gp.seq_type == "0" ? 1 : 2) : 0;
}
gpuPeProcess::~gpuPeProcess() { snk_engine_destroy(d_->eng); delete d_; }

void gpuPeProcess::filter_pe_fqs(PEcalOption* opt)
{
    const int index = (int)(opt->local_fs - local_fs);            // thread_process_reads passes &local_fs[index] (peprocess.cpp:1877)
    const size_t n = opt->fq1s->size();
    if (n != opt->fq2s->size()) { cerr << "Error:reads number in fq1 and fq2 are different" << endl; exit(1); }
    static thread_local Packed p1, p2;
    static thread_local std::vector<snk_read_result> r1, r2;
    p1.pack(*opt->fq1s); p2.pack(*opt->fq2s);
    if (p1.b.stride != p2.b.stride) {                              // the engine wants one stride for both mates
        const size_t want = std::max(p1.b.stride, p2.b.stride);
        for (Packed* p : {&p1, &p2}) {
            if (p->b.stride == want) continue;
            std::vector<uint8_t> s(n * want, 0), q(n * want, 0);
            for (size_t i = 0; i < n; i++) { memcpy(&s[i * want], &p->seq[i * p->b.stride], p->b.stride); memcpy(&q[i * want], &p->qual[i * p->b.stride], p->b.stride); }
            p->seq.swap(s); p->qual.swap(q); p->b.seq = p->seq.data(); p->b.qual = p->qual.data(); p->b.stride = (uint32_t)want;
        }
    }
    r1.resize(n); r2.resize(n);
    uint64_t first;
    {
        std::lock_guard<std::mutex> g(d_->mu);
        first = d_->cursor.first_index(index, n);
        if (!d_->checked) {                                       // the Phred-system sanity check lives in stat_pe_fqs (:1207-1319): run it once on scratch tables
            C_fastq_file_stat t1(gp), t2(gp);
            PEstatOption o; o.fq1s = opt->fq1s; o.fq2s = opt->fq2s; o.stat1 = &t1; o.stat2 = &t2;
            peProcess::stat_pe_fqs(o, "raw");
            d_->checked = true;
        }
    }
    if (snk_filter_pe_host(d_->eng, &p1.b, &p2.b, r1.data(), r2.data(), first)) engine_die();
    const bool want_trim = !gp.trim_fq1.empty(), want_clean = !gp.clean_fq1.empty();
    for (size_t i = 0; i < n; i++) {
        const bool keep = r1[i].category == SNK_KEEP;
        if (!want_trim && !(keep && want_clean)) continue;
        C_fastq a = trimmed((*opt->fq1s)[i], r1[i], d_->id_mode), b = trimmed((*opt->fq2s)[i], r2[i], d_->id_mode);
        if (want_trim) { preOutput(1, a); preOutput(2, b); opt->trim_result1->emplace_back(a); opt->trim_result2->emplace_back(b); }
        if (keep && want_clean) { preOutput(1, a); preOutput(2, b); opt->clean_result1->emplace_back(a); opt->clean_result2->emplace_back(b); }
    }
}

// raw and clean tables were accumulated on the device by the same submission
void* gpuPeProcess::stat_pe_fqs(PEstatOption, string) { return &bq_check; }

void gpuPeProcess::merge_stat()
{
    check_engine_errors(d_->eng);
    std::vector<uint64_t> st((size_t)d_->params.n_slots * snk_stats_slot_words());
    if (snk_engine_stats(d_->eng, st.data())) engine_die();
    for (int i = 0; i < gp.threads_num; i++) {
        const uint64_t* S = st.data() + (size_t)i * SNK_SLOT_WORDS;
        fill_filter_stat(S, local_fs[i]);
        fill_file_stat(S + SNK_SLOT_FILE_OFF(SNK_RAW1), local_raw_stat1[i], gp.maxBaseQuality);
        fill_file_stat(S + SNK_SLOT_FILE_OFF(SNK_RAW2), local_raw_stat2[i], gp.maxBaseQuality);
        fill_file_stat(S + SNK_SLOT_FILE_OFF(SNK_CLEAN1), local_clean_stat1[i], gp.maxBaseQuality);
        fill_file_stat(S + SNK_SLOT_FILE_OFF(SNK_CLEAN2), local_clean_stat2[i], gp.maxBaseQuality);
    }
    peProcess::merge_stat();                                      // the reference's own update_stat, then its print_stat
}

// ================================================================================================ SE
struct gpuSeProcess::Impl {
    snk_engine* eng = nullptr;
    snk_params params;
    SlotCursor cursor;
    int id_mode = 0;
    bool checked = false;
    std::mutex mu;
};

gpuSeProcess::gpuSeProcess(C_global_parameter m_gp) : seProcess(m_gp), d_(new Impl())
{
    const snk::HostParams hp = host_params_of(gp, false);
    snk::to_engine_params(hp, d_->params);
    if (snk_params_check(&d_->params)) engine_die();
    if (snk_engine_create(&d_->params, 0, &d_->eng)) engine_die();
    d_->cursor.T = gp.threads_num; d_->cursor.block = (uint64_t)d_->params.slot_block; d_->cursor.done.assign(gp.threads_num, 0);
    d_->id_mode = gp.index_remove ? (gp.seq_type == "0" ? 1 : 2) : 0;
}
gpuSeProcess::~gpuSeProcess() { snk_engine_destroy(d_->eng); delete d_; }

void gpuSeProcess::filter_se_fqs(SEcalOption opt)
{
    const int index = (int)(opt.se_local_fs - se_local_fs);       // seprocess.cpp:1951
    const size_t n = opt.fq1s->size();
    static thread_local Packed p1;
    static thread_local std::vector<snk_read_result> r1;
    p1.pack(*opt.fq1s);
    r1.resize(n);
    uint64_t first;
    {
        std::lock_guard<std::mutex> g(d_->mu);
        first = d_->cursor.first_index(index, n);
        if (!d_->checked) {
            C_fastq_file_stat t1(gp);
            SEstatOption o; o.fq1s = opt.fq1s; o.stat1 = &t1;
            seProcess::stat_se_fqs(o, "raw");
            d_->checked = true;
        }
    }
    if (snk_filter_se_host(d_->eng, &p1.b, r1.data(), first)) engine_die();
    const bool want_trim = !gp.trim_fq1.empty(), want_clean = !gp.clean_fq1.empty();
    for (size_t i = 0; i < n; i++) {
        const bool keep = r1[i].category == SNK_KEEP;
        if (!want_trim && !(keep && want_clean)) continue;
        C_fastq a = trimmed((*opt.fq1s)[i], r1[i], d_->id_mode);
        if (want_trim) { preOutput(1, a); opt.trim_result1->emplace_back(a); }
        if (keep && want_clean) { preOutput(1, a); opt.clean_result1->emplace_back(a); }
    }
}

void* gpuSeProcess::stat_se_fqs(SEstatOption, string) { return &se_bq_check; }

void gpuSeProcess::merge_stat()
{
    check_engine_errors(d_->eng);
    std::vector<uint64_t> st((size_t)d_->params.n_slots * snk_stats_slot_words());
    if (snk_engine_stats(d_->eng, st.data())) engine_die();
    for (int i = 0; i < gp.threads_num; i++) {
        const uint64_t* S = st.data() + (size_t)i * SNK_SLOT_WORDS;
        fill_filter_stat(S, se_local_fs[i]);
        fill_file_stat(S + SNK_SLOT_FILE_OFF(SNK_RAW1), se_local_raw_stat1[i], gp.maxBaseQuality);
        fill_file_stat(S + SNK_SLOT_FILE_OFF(SNK_CLEAN1), se_local_clean_stat1[i], gp.maxBaseQuality);
    }
    seProcess::merge_stat();
}

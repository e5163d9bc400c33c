// report.cpp — host-side statistics merge + report writer (no GPU code).
//
// Replays the reference's per-thread merge (peProcess::update_stat peprocess.cpp:732-1069,
// seProcess::update_stat seprocess.cpp:436-630) over the engine's per-slot tables and writes the
// report files of peProcess::print_stat (peprocess.cpp:178-731) / seProcess::print_stat
// (seprocess.cpp:96-434) byte for byte, including the partition-dependent truncations, the
// "-nan%" cells, the integer quartile mean (gc.cpp:68-119) and the negative-index spill of the
// trimming tables. A slot is one logical reference thread.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <string>
#include <vector>
#include "../../include/snk_engine.h"
#include "host_common.h"

namespace snk {

namespace {

// Merged view of one FASTQ set: C_fastq_file_stat (global_variable.h:126-135) after merge_stat.
struct Merged {
    uint64_t read_max_length = 0, read_length = 0;
    uint64_t g[SNK_GS_COUNT] = {0};                 // reads, bases, A, C, G, T, N, q20, q30
    std::vector<uint64_t> bs, qs, ts;
    Merged() : bs(SNK_BS_WORDS, 0), qs(SNK_QS_WORDS, 0), ts(SNK_TS_WORDS, 0) {}
    uint64_t base(uint64_t pos, int b) const { return pos < SNK_MAX_READ_LEN ? bs[pos * 5 + b] : 0; }
    uint64_t qual(uint64_t pos, uint64_t q) const { return (pos < SNK_MAX_READ_LEN && q < SNK_QBINS) ? qs[pos * SNK_QBINS + q] : 0; }
    // flat on purpose: row READ_MAX_LEN of one array is row 0 of the next (the reference prints ts.x[read_length])
    uint64_t trim(int arr, uint64_t pos) const { size_t f = (size_t)arr * SNK_MAX_READ_LEN + pos; return f < SNK_TS_WORDS ? ts[f] : 0; }
};

// One slot's view of one FASTQ set (a thread-local C_fastq_file_stat).
struct SlotFile {
    const uint64_t* w;
    explicit SlotFile(const uint64_t* p) : w(p) {}
    uint64_t gs(int i) const { return w[SNK_FILE_GS_OFF + i]; }
    // gs.read_length of a thread = length of the last record it saw (peprocess.cpp:1202,1419)
    uint64_t last_len() const { return w[SNK_FILE_GS_OFF + SNK_GS_LAST_KEY] & 0xFFFFu; }
    const uint64_t* bs() const { return w + SNK_FILE_BS_OFF; }
    const uint64_t* qs() const { return w + SNK_FILE_QS_OFF; }
    const uint64_t* ts() const { return w + SNK_FILE_TS_OFF; }
};

inline uint64_t clamp_rows(uint64_t n) { return n < SNK_MAX_READ_LEN ? n : SNK_MAX_READ_LEN; }
inline uint64_t clamp_q(int64_t q) { return q < 0 ? 0 : (q >= SNK_QBINS ? SNK_QBINS - 1 : (uint64_t)q); }

void add_counts(Merged& m, const SlotFile& s)
{
    for (int i = SNK_GS_READS; i <= SNK_GS_Q30; i++) m.g[i] += s.gs(i);
}
void add_bases(Merged& m, const SlotFile& s, uint64_t rows)
{
    rows = clamp_rows(rows);
    for (uint64_t i = 0; i < rows * 5; i++) m.bs[i] += s.bs()[i];
}
void add_trims(Merged& m, const SlotFile& s, uint64_t lo, uint64_t hi /*exclusive*/)
{
    if (hi > SNK_MAX_READ_LEN) hi = SNK_MAX_READ_LEN;
    for (int a = 0; a < SNK_TS_COUNT; a++)
        for (uint64_t i = lo; i < hi; i++) m.ts[(size_t)a * SNK_MAX_READ_LEN + i] += s.ts()[(size_t)a * SNK_MAX_READ_LEN + i];
}
// highest quality bin in 1..maxBaseQuality that the slot's table uses within `rows`
int slot_max_qual(const SlotFile& s, uint64_t rows, int max_base_quality)
{
    int mq = 0;
    rows = clamp_rows(rows);
    uint64_t top = clamp_q(max_base_quality);
    for (uint64_t i = 0; i < rows; i++)
        for (uint64_t j = 1; j <= top; j++)
            if (s.qs()[i * SNK_QBINS + j] > 0 && (int)j > mq) mq = (int)j;
    return mq;
}
void add_quals(Merged& m, const SlotFile& s, uint64_t rows, int max_qual)
{
    rows = clamp_rows(rows);
    for (uint64_t i = 0; i < rows; i++)
        for (int j = 0; j <= max_qual; j++) m.qs[i * SNK_QBINS + j] += s.qs()[i * SNK_QBINS + j];
}

struct Global {
    uint64_t fs[SNK_FS_COUNT] = {0};
    Merged f[SNK_FILE_COUNT];
};

// merge_stat: for each logical thread, update_stat(raw) then update_stat(clean)
void merge_all(const snk_params& p, const uint64_t* stats, Global& G)
{
    const bool pe = p.is_pe != 0;
    for (int s = 0; s < p.n_slots; s++) {
        const uint64_t* S = stats + (size_t)s * SNK_SLOT_WORDS;
        SlotFile r1(S + SNK_SLOT_FILE_OFF(SNK_RAW1)), r2(S + SNK_SLOT_FILE_OFF(SNK_RAW2));
        SlotFile c1(S + SNK_SLOT_FILE_OFF(SNK_CLEAN1)), c2(S + SNK_SLOT_FILE_OFF(SNK_CLEAN2));
        Merged &R1 = G.f[SNK_RAW1], &R2 = G.f[SNK_RAW2], &C1 = G.f[SNK_CLEAN1], &C2 = G.f[SNK_CLEAN2];
        // ---- raw ----
        if (R1.read_length == 0) R1.read_length = r1.last_len();
        if (R1.read_max_length < r1.last_len()) R1.read_max_length = r1.last_len();
        add_counts(R1, r1);
        if (pe) {
            if (R2.read_length == 0) R2.read_length = r2.last_len();
            if (R2.read_max_length < r2.last_len()) R2.read_max_length = r2.last_len();
            add_counts(R2, r2);
        }
        {
            // every raw loop (both mates) is bounded by raw-1's running read_max_length
            const uint64_t rows = R1.read_max_length;
            add_bases(R1, r1, rows);
            if (pe) add_bases(R2, r2, rows);
            if (pe) { add_trims(R1, r1, 0, rows); add_trims(R2, r2, 0, rows); }
            else add_trims(R1, r1, 1, rows + 1);            // seprocess.cpp:464 runs 1..max inclusive
            const int mq = slot_max_qual(r1, rows, p.max_base_quality);   // fq1's table decides for both mates
            add_quals(R1, r1, rows, mq);
            if (pe) add_quals(R2, r2, rows, mq);
        }
        for (int i = 0; i < SNK_FS_COUNT; i++) G.fs[i] += S[i];
        // ---- clean ----
        add_counts(C1, c1);
        C1.read_length = (C1.g[SNK_GS_READS] == 0) ? c1.last_len() : C1.g[SNK_GS_BASES] / C1.g[SNK_GS_READS];
        if (C1.read_max_length < c1.last_len()) C1.read_max_length = c1.last_len();
        if (pe) {
            add_counts(C2, c2);
            C2.read_length = (C2.g[SNK_GS_READS] == 0) ? c2.last_len() : C2.g[SNK_GS_BASES] / C2.g[SNK_GS_READS];
            if (C2.read_max_length < C2.read_length) C2.read_max_length = C2.read_length;   // own integer mean (peprocess.cpp:992)
        }
        add_bases(C1, c1, C1.read_max_length);
        add_trims(C1, c1, 0, C1.read_max_length);
        add_quals(C1, c1, C1.read_max_length, slot_max_qual(c1, C1.read_max_length, p.max_base_quality));
        if (pe) {
            add_bases(C2, c2, C2.read_max_length);
            add_trims(C2, c2, 0, C2.read_max_length);
            add_quals(C2, c2, C2.read_max_length, slot_max_qual(c2, C2.read_max_length, p.max_base_quality));
        }
    }
}

// gc.cpp:68-119 cal_quar_from_array, with its int accumulators (wrap like the x86 build)
struct Quartiles { float mean, median, lower, upper, p10, p90; };
Quartiles quartiles(const Merged& m, uint64_t pos, int len)
{
    Quartiles r;
    r.mean = r.median = r.lower = r.upper = r.p10 = r.p90 = 0;
    unsigned long long total = 0;
    int32_t n = 0;
    for (int i = 0; i <= len; i++) {
        uint64_t d = m.qual(pos, (uint64_t)i);
        total += (unsigned long long)i * d;
        n = (int32_t)((uint32_t)n + (uint32_t)d);
    }
    r.mean = (n == 0) ? 0 : (float)(total / (unsigned long long)(long long)n);
    const int32_t lower_pos = n / 4, upper_pos = (int32_t)((uint32_t)n * 3u) / 4, p10_pos = n / 10,
                  p90_pos = (int32_t)((uint32_t)n * 9u) / 10, med_pos = n / 2;
    int32_t last = 0, cur = 0;
    for (int i = 0; i <= len; i++) {
        cur = (int32_t)((uint32_t)cur + (uint32_t)m.qual(pos, (uint64_t)i));
        if (lower_pos >= last && lower_pos <= cur) r.lower = (float)i;
        if (upper_pos >= last && upper_pos <= cur) r.upper = (float)i;
        if (p10_pos >= last && p10_pos <= cur) r.p10 = (float)i;
        if (p90_pos >= last && p90_pos <= cur) r.p90 = (float)i;
        if (med_pos >= last && med_pos <= cur) r.median = (float)i;
        last = cur;
    }
    return r;
}

struct FilterRow { const char* label; int fs_base; };
// label order of peprocess.cpp:226-241 restricted to the categories the engine produces
const FilterRow kPeRows[] = {
    {"Reads with filtered tile", SNK_FS_TILE}, {"Reads with filtered fov", SNK_FS_FOV},
    {"Reads too short", SNK_FS_SHORT}, {"Reads too long", SNK_FS_LONG},
    {"Reads with global contam sequence", SNK_FS_GCONTAM}, {"Reads with contam sequence", SNK_FS_CONTAM},
    {"Reads with n rate exceed", SNK_FS_N}, {"Reads with highA", SNK_FS_HIGHA},
    {"Reads with polyX", SNK_FS_POLYX}, {"Reads with low quality", SNK_FS_LOWQ},
    {"Reads with low mean quality", SNK_FS_MEANQ}, {"Reads with adapter", SNK_FS_ADAPTER}};
// seprocess.cpp:136-150: the global contaminants come last there
const FilterRow kSeRows[] = {
    {"Reads with filtered tile", SNK_FS_TILE}, {"Reads with filtered fov", SNK_FS_FOV},
    {"Reads too short", SNK_FS_SHORT}, {"Reads too long", SNK_FS_LONG}, {"Reads with contam sequence", SNK_FS_CONTAM},
    {"Reads with n rate exceed", SNK_FS_N}, {"Reads with highA", SNK_FS_HIGHA},
    {"Reads with polyX", SNK_FS_POLYX}, {"Reads with low quality", SNK_FS_LOWQ},
    {"Reads with low mean quality", SNK_FS_MEANQ}, {"Reads with adapter", SNK_FS_ADAPTER},
    {"Reads with global contam sequence", SNK_FS_GCONTAM}};
const int kRows = 12;

uint64_t filtered_total(const Global& G)
{
    uint64_t t = 0;
    for (int i = 0; i < kRows; i++) t += G.fs[kPeRows[i].fs_base];
    return t;
}

std::string pct2(float v)
{
    char buf[100];
    snprintf(buf, sizeof buf, "%.2f", v);
    return buf;
}
// the seven "xx.xx" cells of one column of Basic_Statistics (peprocess.cpp:341-347)
void ratio_cells(const Merged& m, std::string out[7])
{
    if (m.g[SNK_GS_READS] == 0) { for (int i = 0; i < 7; i++) out[i] = ""; return; }   // reference leaves the buffers untouched
    const int idx[7] = {SNK_GS_A, SNK_GS_C, SNK_GS_G, SNK_GS_T, SNK_GS_N, SNK_GS_Q20, SNK_GS_Q30};
    for (int i = 0; i < 7; i++) out[i] = pct2(100 * (float)m.g[idx[i]] / m.g[SNK_GS_BASES]);
}

void write_trim_file(std::ofstream& of, const Merged& raw, const Merged& clean, uint64_t read_length, bool se)
{
    of << "Pos\tHeadLowQual\tHeadFixLen\tTailAdapter\tTailLowQual\tTailFixLen\tCleanHeadLowQual\tCleanHeadFixLen\tCleanTailAdapter\tCleanTailLowQual\tCleanTailFixLen" << std::endl;
    // totals over rows 0..read_length-1, rows printed 1..read_length (peprocess.cpp:607-620)
    uint64_t head_raw = 0, tail_raw = 0, head_clean = 0, tail_clean = 0;
    for (uint64_t i = 0; i < read_length; i++) {
        head_raw += raw.trim(SNK_TS_HT, i) + raw.trim(SNK_TS_HLQ, i);
        tail_raw += raw.trim(SNK_TS_TA, i) + raw.trim(SNK_TS_TLQ, i) + raw.trim(SNK_TS_TT, i);
        head_clean += clean.trim(SNK_TS_HT, i) + clean.trim(SNK_TS_HLQ, i);
        tail_clean += clean.trim(SNK_TS_TA, i) + clean.trim(SNK_TS_TLQ, i) + clean.trim(SNK_TS_TT, i);
    }
    (void)se;
    auto cell = [&](uint64_t v, uint64_t total, const char* sep) {
        of << v << "\t" << std::setiosflags(std::ios::fixed) << std::setprecision(2) << 100 * (float)v / total << sep;
    };
    auto zero = [&](uint64_t v, const char* sep) { of << v << "\t0.00" << sep; };
    for (uint64_t i = 1; i <= read_length; i++) {
        of << i << "\t";
        if (head_raw > 0) { cell(raw.trim(SNK_TS_HLQ, i), head_raw, "%\t"); cell(raw.trim(SNK_TS_HT, i), head_raw, "%\t"); }
        else { zero(raw.trim(SNK_TS_HLQ, i), "%\t"); zero(raw.trim(SNK_TS_HT, i), "%\t"); }
        if (tail_raw > 0) { cell(raw.trim(SNK_TS_TA, i), tail_raw, "%\t"); cell(raw.trim(SNK_TS_TLQ, i), tail_raw, "%\t"); cell(raw.trim(SNK_TS_TT, i), tail_raw, "%\t"); }
        else { zero(raw.trim(SNK_TS_TA, i), "%\t"); zero(raw.trim(SNK_TS_TLQ, i), "%\t"); zero(raw.trim(SNK_TS_TLQ, i), "%\t"); }   // tlq twice, as the reference
        if (head_clean > 0) { cell(clean.trim(SNK_TS_HLQ, i), head_clean, "%\t"); cell(clean.trim(SNK_TS_HT, i), head_clean, "%\t"); }
        else { zero(clean.trim(SNK_TS_HLQ, i), "%\t"); zero(clean.trim(SNK_TS_HT, i), "%\t"); }
        if (tail_clean > 0) {
            cell(clean.trim(SNK_TS_TA, i), tail_clean, "%\t"); cell(clean.trim(SNK_TS_TLQ, i), tail_clean, "%\t");
            of << clean.trim(SNK_TS_TT, i) << "\t" << std::setiosflags(std::ios::fixed) << std::setprecision(2)
               << 100 * (float)clean.trim(SNK_TS_TT, i) / tail_clean << "%" << std::endl;
        } else {
            zero(clean.trim(SNK_TS_TA, i), "%\t"); zero(clean.trim(SNK_TS_TLQ, i), "%\t");
            of << clean.trim(SNK_TS_TLQ, i) << "\t0.00%" << std::endl;
        }
    }
}

void write_base_file(std::ofstream& of, const Merged& raw, const Merged& clean, uint64_t rows)
{
    of << "Pos\tA\tC\tG\tT\tN\tclean A\tclean C\tclean G\tclean T\tclean N" << std::endl;
    for (uint64_t i = 0; i < rows; i++) {
        of << i + 1 << "\t";
        float raw_total = 0, clean_total = 0;       // float accumulators, as the reference
        for (int j = 0; j < 5; j++) { raw_total += raw.base(i, j); clean_total += clean.base(i, j); }
        for (int j = 0; j < 5; j++)
            of << std::setiosflags(std::ios::fixed) << std::setprecision(2) << 100 * (float)raw.base(i, j) / raw_total << "%\t";
        for (int j = 0; j < 5; j++) {
            of << std::setiosflags(std::ios::fixed) << std::setprecision(2) << 100 * (float)clean.base(i, j) / clean_total << "%";
            if (j != 4) of << "\t"; else of << std::endl;
        }
    }
}

int print_max_qual(const Merged& raw1, int max_base_quality)
{
    int mq = 0;
    uint64_t top = clamp_q(max_base_quality);
    for (uint64_t i = 0; i < raw1.read_length; i++)
        for (uint64_t j = 1; j <= top; j++)
            if (raw1.qual(i, j) > 0 && (int)j > mq) mq = (int)j;
    return mq;
}

void qual_header(std::ofstream& of, int max_qual)
{
    of << "Pos\t";
    for (int i = 0; i <= max_qual; i++) of << "Q" << i << "\t";
    of << "Mean\tMedian\tLower quartile\tUpper quartile\t10th percentile\t90th percentile" << std::endl;
}
// one row of the quality table; returns (q20 fraction, q30 fraction) of that position
void qual_row(std::ofstream& of, const Merged& m, uint64_t pos, int max_qual, int quart_len, float& q20, float& q30)
{
    of << pos + 1 << "\t";
    uint64_t n20 = 0, n30 = 0, total = 0;
    for (int j = 0; j <= max_qual; j++) {
        uint64_t v = m.qual(pos, (uint64_t)j);
        if (j >= 20) n20 += v;
        if (j >= 30) n30 += v;
        total += v;
        of << std::setiosflags(std::ios::fixed);
        of << std::setprecision(0) << v << "\t";
    }
    q20 = (float)n20 / total;
    q30 = (float)n30 / total;
    Quartiles q = quartiles(m, pos, quart_len);
    of << std::setiosflags(std::ios::fixed) << std::setprecision(2) << q.mean << "\t";
    of << std::setprecision(0) << q.median << "\t" << q.lower << "\t" << q.upper << "\t" << q.p10 << "\t" << q.p90 << std::endl;
}

bool open_out(std::ofstream& of, const std::string& path)
{
    of.open(path.c_str());
    if (!of) { set_error("cannot open such file," + path); return false; }
    return true;
}

} // namespace

int report_write_pe(const snk_params& p, const uint64_t* stats, const std::string& dir)
{
    Global* Gp = new Global();
    Global& G = *Gp;
    merge_all(p, stats, G);
    const Merged &R1 = G.f[SNK_RAW1], &R2 = G.f[SNK_RAW2], &C1 = G.f[SNK_CLEAN1], &C2 = G.f[SNK_CLEAN2];
    std::ofstream f_filter, f_general, f_bs1, f_bs2, f_qs1, f_qs2, f_q1, f_q2, f_t1, f_t2;
    if (!open_out(f_filter, dir + "/Statistics_of_Filtered_Reads.txt") ||
        !open_out(f_general, dir + "/Basic_Statistics_of_Sequencing_Quality.txt") ||
        !open_out(f_bs1, dir + "/Base_distributions_by_read_position_1.txt") ||
        !open_out(f_bs2, dir + "/Base_distributions_by_read_position_2.txt") ||
        !open_out(f_qs1, dir + "/Base_quality_value_distribution_by_read_position_1.txt") ||
        !open_out(f_qs2, dir + "/Base_quality_value_distribution_by_read_position_2.txt") ||
        !open_out(f_q1, dir + "/Distribution_of_Q20_Q30_bases_by_read_position_1.txt") ||
        !open_out(f_q2, dir + "/Distribution_of_Q20_Q30_bases_by_read_position_2.txt") ||
        !open_out(f_t1, dir + "/Statistics_of_Trimming_Position_of_Reads_1.txt") ||
        !open_out(f_t2, dir + "/Statistics_of_Trimming_Position_of_Reads_2.txt")) { delete Gp; return 1; }

    // ---- Statistics_of_Filtered_Reads.txt (peprocess.cpp:225-322) ----
    const uint64_t total = filtered_total(G);
    f_filter << "Item\t\t\t\tTotal\tPercentage\tfastq1\tfastq2\toverlap" << std::endl;
    f_filter << std::setiosflags(std::ios::fixed);
    f_filter << "Total filtered read pair number\t" << total << "\t100.00%\t\t" << total << "\t" << total << "\t" << total << std::endl;
    for (int i = 0; i < kRows; i++) {
        const uint64_t* c = G.fs + kPeRows[i].fs_base;
        if (c[0] == 0) continue;
        f_filter << kPeRows[i].label << "\t" << c[0] << "\t";
        f_filter << std::setprecision(2) << 100 * (float)c[0] / total << "%\t";
        // tile / fov have no per-mate counters: the reference prints the same number four times (peprocess.cpp:271-304)
        const bool plain = kPeRows[i].fs_base == SNK_FS_TILE || kPeRows[i].fs_base == SNK_FS_FOV;
        f_filter << (plain ? c[0] : c[1]) << "\t" << (plain ? c[0] : c[2]) << "\t" << (plain ? c[0] : c[3]) << std::endl;
    }
    f_filter.close();

    // ---- Basic_Statistics_of_Sequencing_Quality.txt (peprocess.cpp:324-413) ----
    f_general << "Item\traw reads(fq1)\tclean reads(fq1)\traw reads(fq2)\tclean reads(fq2)" << std::endl;
    const Merged* col[4] = {&R1, &C1, &R2, &C2};
    float rl[4] = {0, 0, 0, 0};
    std::string cells[4][7];
    std::string filt_ratio[2];
    for (int k = 0; k < 4; k++) {
        if (col[k]->g[SNK_GS_READS] != 0) rl[k] = 1.0 * col[k]->g[SNK_GS_BASES] / col[k]->g[SNK_GS_READS];
        ratio_cells(*col[k], cells[k]);
    }
    if (R1.g[SNK_GS_READS] != 0) filt_ratio[0] = pct2(100 * (float)total / R1.g[SNK_GS_READS]);
    if (R2.g[SNK_GS_READS] != 0) filt_ratio[1] = pct2(100 * (float)total / R2.g[SNK_GS_READS]);
    f_general << std::setiosflags(std::ios::fixed) << std::setprecision(1) << "Read length\t" << rl[0] << "\t" << rl[1] << "\t" << rl[2] << "\t" << rl[3] << std::endl;
    f_general << "Total number of reads\t" << std::setprecision(15) << R1.g[SNK_GS_READS] << " (100.00%)\t" << C1.g[SNK_GS_READS]
              << " (100.00%)\t" << R2.g[SNK_GS_READS] << " (100.00%)\t" << C2.g[SNK_GS_READS] << " (100.00%)" << std::endl;
    f_general << "Number of filtered reads\t" << total << " (" << filt_ratio[0] << "%)\t-\t" << total << " (" << filt_ratio[1] << "%)\t-" << std::endl;
    f_general << "Total number of bases\t" << std::setprecision(15) << R1.g[SNK_GS_BASES] << " (100.00%)\t" << C1.g[SNK_GS_BASES]
              << " (100.00%)\t" << R2.g[SNK_GS_BASES] << " (100.00%)\t" << C2.g[SNK_GS_BASES] << " (100.00%)" << std::endl;
    const uint64_t filt_bases = total * R1.read_length;      // raw-1 read_length for both mates (peprocess.cpp:387-388)
    f_general << "Number of filtered bases\t" << std::setprecision(15) << filt_bases << " (" << filt_ratio[0] << "%)\t-\t" << filt_bases
              << " (" << filt_ratio[1] << "%)\t-" << std::endl;
    const char* names[7] = {"Number of base A", "Number of base C", "Number of base G", "Number of base T", "Number of base N", "Q20 number", "Q30 number"};
    const int gidx[7] = {SNK_GS_A, SNK_GS_C, SNK_GS_G, SNK_GS_T, SNK_GS_N, SNK_GS_Q20, SNK_GS_Q30};
    for (int r = 0; r < 7; r++) {
        f_general << names[r] << "\t" << std::setprecision(15);
        for (int k = 0; k < 4; k++) {
            f_general << col[k]->g[gidx[r]] << " (" << cells[k][r] << "%)";
            if (k != 3) f_general << "\t";
        }
        f_general << std::endl;
    }
    f_general.close();

    // ---- Base_distributions_by_read_position_{1,2}.txt (peprocess.cpp:414-466) ----
    write_base_file(f_bs1, R1, C1, R1.read_length);
    write_base_file(f_bs2, R2, C2, R1.read_length);
    f_bs1.close(); f_bs2.close();

    // ---- Base_quality_value_distribution + Q20/Q30 (peprocess.cpp:468-602) ----
    const int max_qual = print_max_qual(R1, p.max_base_quality);
    const uint64_t rows = R1.read_max_length > R2.read_max_length ? R1.read_max_length : R2.read_max_length;
    f_qs1 << "#raw fastq1 quality distribution" << std::endl;
    f_qs2 << "#raw fastq2 quality distribution" << std::endl;
    qual_header(f_qs1, max_qual); qual_header(f_qs2, max_qual);
    std::vector<float> r1q20(rows), r1q30(rows), r2q20(rows), r2q30(rows);
    for (uint64_t i = 0; i < rows; i++) {
        qual_row(f_qs1, R1, i, max_qual, max_qual, r1q20[i], r1q30[i]);
        qual_row(f_qs2, R2, i, max_qual, max_qual, r2q20[i], r2q30[i]);
    }
    f_qs1 << "#clean fastq1 quality distribution" << std::endl;
    f_qs2 << "#clean fastq2 quality distribution" << std::endl;
    qual_header(f_qs1, max_qual); qual_header(f_qs2, max_qual);
    const char* qhdr = "Position in reads\tPercentage of Q20+ bases\tPercentage of Q30+ bases\tPercentage of Clean Q20+\tPercentage of Clean Q30+";
    f_q1 << qhdr << std::endl;
    f_q2 << qhdr << std::endl;
    for (uint64_t i = 0; i < rows; i++) {
        float c1q20, c1q30, c2q20, c2q30;
        qual_row(f_qs1, C1, i, max_qual, max_qual, c1q20, c1q30);
        qual_row(f_qs2, C2, i, max_qual, max_qual, c2q20, c2q30);
        f_q1 << i + 1 << std::setiosflags(std::ios::fixed) << std::setprecision(2) << "\t" << 100 * r1q20[i] << "%\t" << 100 * r1q30[i]
             << "%\t" << 100 * c1q20 << "%\t" << 100 * c1q30 << "%" << std::endl;
        f_q2 << i + 1 << std::setiosflags(std::ios::fixed) << std::setprecision(2) << "\t" << 100 * r2q20[i] << "%\t" << 100 * r2q30[i]
             << "%\t" << 100 * c2q20 << "%\t" << 100 * c2q30 << "%" << std::endl;
    }
    f_qs1.close(); f_qs2.close(); f_q1.close(); f_q2.close();

    // ---- Statistics_of_Trimming_Position_of_Reads_{1,2}.txt (peprocess.cpp:603-715) ----
    write_trim_file(f_t1, R1, C1, R1.read_length, false);
    write_trim_file(f_t2, R2, C2, R1.read_length, false);
    f_t1.close(); f_t2.close();
    delete Gp;
    return 0;
}

int report_write_se(const snk_params& p, const uint64_t* stats, const std::string& dir)
{
    Global* Gp = new Global();
    Global& G = *Gp;
    merge_all(p, stats, G);
    const Merged &R1 = G.f[SNK_RAW1], &C1 = G.f[SNK_CLEAN1];
    std::ofstream f_filter, f_general, f_bs1, f_qs1, f_q1, f_t1;
    if (!open_out(f_filter, dir + "/Statistics_of_Filtered_Reads.txt") ||
        !open_out(f_general, dir + "/Basic_Statistics_of_Sequencing_Quality.txt") ||
        !open_out(f_bs1, dir + "/Base_distributions_by_read_position_1.txt") ||
        !open_out(f_qs1, dir + "/Base_quality_value_distribution_by_read_position_1.txt") ||
        !open_out(f_q1, dir + "/Distribution_of_Q20_Q30_bases_by_read_position_1.txt") ||
        !open_out(f_t1, dir + "/Statistics_of_Trimming_Position_of_Reads_1.txt")) { delete Gp; return 1; }

    // seprocess.cpp:135-181. The SE update_stat merges tile_num but not fov_num (seprocess.cpp:495): reads removed
    // by the fov list never reach the report (neither their row nor the totals)
    G.fs[SNK_FS_FOV] = 0;
    const uint64_t total = filtered_total(G);
    f_filter << "Item\tTotal\tPercentage" << std::endl;
    f_filter << std::setiosflags(std::ios::fixed);
    f_filter << "Total filtered read pair number\t" << total << "\t100.00%" << std::endl;
    for (int i = 0; i < kRows; i++) {
        const uint64_t c = G.fs[kSeRows[i].fs_base];
        if (c == 0) continue;
        f_filter << kSeRows[i].label << "\t" << c << "\t";
        f_filter << std::setprecision(2) << 100 * (float)c / total << "%" << std::endl;
    }
    f_filter.close();

    // seprocess.cpp:182-235
    f_general << "Item\traw reads(fq1)\tclean reads(fq1)" << std::endl;
    float raw_rl = 0, clean_rl = 0;
    std::string rc[7], cc[7], filt_ratio;
    if (R1.g[SNK_GS_READS] != 0) { raw_rl = (float)R1.g[SNK_GS_BASES] / R1.g[SNK_GS_READS]; filt_ratio = pct2(100 * (float)total / R1.g[SNK_GS_READS]); }
    if (C1.g[SNK_GS_READS] != 0) clean_rl = (float)C1.g[SNK_GS_BASES] / C1.g[SNK_GS_READS];
    ratio_cells(R1, rc); ratio_cells(C1, cc);
    f_general << std::setiosflags(std::ios::fixed) << std::setprecision(1) << "Read length\t" << raw_rl << "\t" << clean_rl << std::endl;
    f_general << "Total number of reads\t" << std::setprecision(15) << R1.g[SNK_GS_READS] << " (100.00%)\t" << C1.g[SNK_GS_READS] << " (100.00%)" << std::endl;
    f_general << "Number of filtered reads\t" << total << " (" << filt_ratio << "%)\t-" << std::endl;
    const uint64_t filt_bases = total * R1.read_length;
    f_general << "Total number of bases\t" << std::setprecision(15) << R1.g[SNK_GS_BASES] << " (100.00%)\t" << C1.g[SNK_GS_BASES] << " (100.00%)" << std::endl;
    f_general << "Number of filtered bases\t" << std::setprecision(15) << filt_bases << " (" << filt_ratio << "%)\t-" << std::endl;
    const char* names[7] = {"Number of base A", "Number of base C", "Number of base G", "Number of base T", "Number of base N", "Q20 number", "Q30 number"};
    const int gidx[7] = {SNK_GS_A, SNK_GS_C, SNK_GS_G, SNK_GS_T, SNK_GS_N, SNK_GS_Q20, SNK_GS_Q30};
    for (int r = 0; r < 7; r++) {
        f_general << names[r] << "\t" << std::setprecision(15) << R1.g[gidx[r]] << " (" << rc[r] << "%)\t" << C1.g[gidx[r]] << " (" << cc[r] << "%)";
        if (r < 5) f_general << "\t";            // base lines carry a trailing tab (seprocess.cpp:219-228)
        f_general << std::endl;
    }
    f_general.close();

    write_base_file(f_bs1, R1, C1, (uint64_t)(int)R1.read_length);
    f_bs1.close();

    // seprocess.cpp:270-361
    const int max_qual = print_max_qual(R1, p.max_base_quality);
    f_qs1 << "#raw fastq1 quality distribution" << std::endl;
    qual_header(f_qs1, max_qual);
    std::vector<float> rq20(R1.read_max_length > C1.read_max_length ? R1.read_max_length : C1.read_max_length, 0.0f), rq30(rq20.size(), 0.0f);
    for (uint64_t i = 0; i < R1.read_length; i++) qual_row(f_qs1, R1, i, max_qual, max_qual + 1, rq20[i], rq30[i]);
    f_qs1 << "#clean fastq1 quality distribution" << std::endl;
    qual_header(f_qs1, max_qual);
    f_q1 << "Position in reads\tPercentage of Q20+ bases\tPercentage of Q30+ bases\tPercentage of Clean Q20+\tPercentage of Clean Q30+" << std::endl;
    for (uint64_t i = 0; i < C1.read_max_length; i++) {
        float cq20, cq30;
        qual_row(f_qs1, C1, i, max_qual, max_qual + 1, cq20, cq30);
        f_q1 << i + 1 << std::setiosflags(std::ios::fixed) << std::setprecision(4) << "\t" << rq20[i] << "\t" << rq30[i] << "\t" << cq20 << "\t" << cq30 << std::endl;
    }
    f_qs1.close(); f_q1.close();

    write_trim_file(f_t1, R1, C1, (uint64_t)(int)R1.read_length, true);
    f_t1.close();
    delete Gp;
    return 0;
}

} // namespace snk

extern "C" {
int snk_report_write_pe(const snk_params* p, const uint64_t* stats, const char* out_dir)
{
    if (!p || !stats || !out_dir) { snk::set_error("null argument"); return 1; }
    return snk::report_write_pe(*p, stats, out_dir);
}
int snk_report_write_se(const snk_params* p, const uint64_t* stats, const char* out_dir)
{
    if (!p || !stats || !out_dir) { snk::set_error("null argument"); return 1; }
    return snk::report_write_se(*p, stats, out_dir);
}
}

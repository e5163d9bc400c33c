// process.cpp — the pinned-buffer batching driver behind peProcess::process / seProcess::process.
// See process.h for the stage diagram. Reference behaviour reproduced here (file:line):
//   line handling of sub_thread           peprocess.cpp:2066-2076, 2090-2131 (.gz: strip the first line's
//                                         trailing-whitespace count from every line), :2198-2239 (plain: strip 1)
//   first-batch pair-ID / Phred checks    peprocess.cpp:1884-1908, 1207-1319; seprocess.cpp:741-867
//   record formatting                     peprocess.cpp:3383-3433 (output_fastqs), :1617-1629 (preOutput /1 /2),
//                                         read_filter.cpp:357-382 (index removal)
//   emission order of the clean records   peprocess.cpp:2141,2248,2957-2990 (see Writer::route)
//   gzip level 2 members                  peprocess.cpp:1803-1810
#include "process.h"
#include "host_common.h"
#include <zlib.h>
#include <sys/stat.h>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

namespace snk {

namespace {

[[noreturn]] void die(const std::string& msg)
{
    std::cerr << "Error:" << msg << std::endl;
    exit(1);
}
void engine_check(int rc) { if (rc) die(snk_last_error()); }

double now_s()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

std::string local_time()
{
    time_t t = time(nullptr);
    char buf[64];
    strftime(buf, sizeof buf, "%Y-%m-%d %H:%M:%S", localtime(&t));
    return buf;
}

void mkdir_p(const std::string& dir)
{
    std::string cur;
    for (size_t i = 0; i <= dir.size(); i++) {
        if (i == dir.size() || dir[i] == '/') {
            if (!cur.empty()) mkdir(cur.c_str(), 0755);
        }
        if (i < dir.size()) cur += dir[i];
    }
}

// ------------------------------------------------------------------ blocking queue
template <typename T>
class Channel {
public:
    void push(T v) { { std::lock_guard<std::mutex> g(m_); q_.push_back(v); } cv_.notify_one(); }
    bool pop(T& v)
    {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        v = q_.front(); q_.pop_front();
        return true;
    }
    void close() { { std::lock_guard<std::mutex> g(m_); closed_ = true; } cv_.notify_all(); }
private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<T> q_;
    bool closed_ = false;
};

// ------------------------------------------------------------------ line reader (gz or plain through zlib)
class LineReader {
public:
    explicit LineReader(const std::string& path) : path_(path)
    {
        f_ = gzopen(path.c_str(), "rb");
        if (!f_) die("cannot open the file," + path);
        gzbuffer(f_, 1 << 22);
        buf_.resize(1 << 24);
    }
    ~LineReader() { if (f_) gzclose(f_); }
    // next line including its '\n' when present; false at EOF
    bool next(const char*& p, size_t& n)
    {
        for (;;) {
            const char* nl = (const char*)memchr(buf_.data() + pos_, '\n', end_ - pos_);
            if (nl) { p = buf_.data() + pos_; n = (size_t)(nl - p) + 1; pos_ += n; return true; }
            if (eof_) {
                if (pos_ < end_) { p = buf_.data() + pos_; n = end_ - pos_; pos_ = end_; return true; }
                return false;
            }
            refill();
        }
    }
private:
    void refill()
    {
        if (pos_ > 0) { memmove(buf_.data(), buf_.data() + pos_, end_ - pos_); end_ -= pos_; pos_ = 0; }
        if (end_ == buf_.size()) buf_.resize(buf_.size() * 2);
        int got = gzread(f_, buf_.data() + end_, (unsigned)std::min<size_t>(buf_.size() - end_, 1u << 30));
        if (got < 0) die("cannot read the file," + path_);
        if (got == 0) eof_ = true;
        end_ += (size_t)got;
    }
    std::string path_;
    gzFile f_ = nullptr;
    std::vector<char> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false;
};

// ------------------------------------------------------------------ one batch travelling through the stages
struct Piece { int kind; std::string bytes; };   // kind: 0 main, 1 deferred, 2 flush-deferred marker
struct MateBuf {
    uint8_t* seq = nullptr; uint8_t* qual = nullptr; uint16_t* len = nullptr; snk_read_result* res = nullptr;
    size_t cap_reads = 0, stride = 0;
    std::vector<char> ids; std::vector<uint32_t> id_off;
    std::vector<Piece> out;
    void release()
    {
        if (seq) snk_host_free(seq);
        if (qual) snk_host_free(qual);
        if (len) snk_host_free(len);
        if (res) snk_host_free(res);
        seq = qual = nullptr; len = nullptr; res = nullptr;
    }
    void reserve(size_t reads, size_t new_stride)
    {
        if (reads <= cap_reads && new_stride == stride) return;
        uint8_t *ns = nullptr, *nq = nullptr; uint16_t* nl = nullptr; snk_read_result* nr = nullptr;
        engine_check(snk_host_alloc((void**)&ns, reads * new_stride + 64));
        engine_check(snk_host_alloc((void**)&nq, reads * new_stride + 64));
        engine_check(snk_host_alloc((void**)&nl, reads * sizeof(uint16_t) + 64));
        engine_check(snk_host_alloc((void**)&nr, reads * sizeof(snk_read_result) + 64));
        memset(ns, 0, reads * new_stride + 64); memset(nq, 0, reads * new_stride + 64);
        if (seq) {      // re-stride what is already there (a longer read appeared: rare)
            const size_t keep = std::min(cap_reads, reads);
            for (size_t i = 0; i < keep; i++) {
                memcpy(ns + i * new_stride, seq + i * stride, std::min(stride, new_stride));
                memcpy(nq + i * new_stride, qual + i * stride, std::min(stride, new_stride));
            }
            memcpy(nl, len, keep * sizeof(uint16_t));
        }
        release();
        seq = ns; qual = nq; len = nl; res = nr; cap_reads = reads; stride = new_stride;
    }
};
struct HostBatch {
    uint64_t seq_no = 0, first_index = 0;
    uint32_t n = 0;
    int gpu = 0, lane = 0;
    MateBuf m[2];
};

inline size_t round16(size_t v) { return (v + 15) / 16 * 16; }

} // namespace

// ==================================================================== FilterRun
class FilterRun {
public:
    FilterRun(const HostParams& hp, bool pe) : hp_(hp), pe_(pe), mates_(pe ? 2 : 1) { to_engine_params(hp_, ep_); ep_.is_pe = pe; }
    void process();

private:
    HostParams hp_;
    bool pe_;
    int mates_;
    snk_params ep_;
    std::vector<snk_engine*> engines_;
    std::ofstream log_;
    std::mutex log_mu_;
    int strip_gz_ = 1;                 // spaceNum of the first line (peprocess.cpp:2066-2076)
    size_t stride_ = 0;                // current row stride (grows when a longer read shows up)
    uint64_t total_reads_ = 0;
    std::string pending_deferred_[2];  // deferred output (already encoded) waiting for its insertion point

    std::vector<HostBatch> batches_;
    Channel<HostBatch*> free_q_, gpu_q_, fmt_q_;
    std::mutex done_mu_;
    std::condition_variable done_cv_;
    std::map<uint64_t, HostBatch*> done_;       // formatted batches by sequence number
    bool fmt_finished_ = false;
    uint64_t n_batches_total_ = 0;

    // stages
    void ingest();
    size_t parse_mate(LineReader& r, HostBatch& b, int mate, size_t max_reads);
    void first_batch_checks(const HostBatch& b);
    void gpu_stage();
    void format_worker();
    void format_mate(HostBatch& b, int mate);
    void writer();
    void encode(std::string& text);
    size_t inflight_depth() const { return engines_.size() * (size_t)snk_engine_lanes(engines_[0]); }
    void log_line(const std::string& s) { std::lock_guard<std::mutex> g(log_mu_); log_ << s << std::endl; }

    // output routing (reference emission order)
    uint64_t cyc_ = 0, defer_len_ = 0, insert_off_ = 0;
    bool reorder_ = false;
    // busy seconds per stage (log only)
    double t_parse_ = 0, t_gpu_wait_ = 0, t_write_ = 0, t_setup_ = 0;
    std::atomic<uint64_t> t_format_us_{0};
};

// ---- parse up to max_reads records of one mate into the pinned SoA rows
size_t FilterRun::parse_mate(LineReader& r, HostBatch& b, int mate, size_t max_reads)
{
    MateBuf& mb = b.m[mate];
    mb.ids.clear(); mb.id_off.clear(); mb.id_off.push_back(0);
    mb.reserve(max_reads, stride_ ? stride_ : 160);
    const size_t strip = hp_.input_gz ? (size_t)strip_gz_ : 1;     // plain: erase(size()-1) (peprocess.cpp:2206)
    size_t n = 0;
    const char* p; size_t ln;
    while (n < max_reads) {
        if (!r.next(p, ln)) break;                                       // id line
        size_t idn = ln > strip ? ln - strip : 0;
        mb.ids.insert(mb.ids.end(), p, p + idn);
        mb.id_off.push_back((uint32_t)mb.ids.size());
        const char* sp; size_t sn;
        if (!r.next(sp, sn)) die("input fastq is truncated," + (mate ? hp_.fq2_path : hp_.fq1_path));
        sn = sn > strip ? sn - strip : 0;
        if (sn > SNK_MAX_READ_LEN) die("read longer than 1000 bases is not supported (READ_MAX_LEN)");
        if (sn > mb.stride) mb.reserve(mb.cap_reads, round16(sn));
        memcpy(mb.seq + n * mb.stride, sp, sn);
        if (sn < mb.stride) memset(mb.seq + n * mb.stride + sn, 0, mb.stride - sn);
        mb.len[n] = (uint16_t)sn;
        if (!r.next(p, ln)) die("input fastq is truncated," + (mate ? hp_.fq2_path : hp_.fq1_path));   // '+'
        const char* qp; size_t qn;
        if (!r.next(qp, qn)) die("input fastq is truncated," + (mate ? hp_.fq2_path : hp_.fq1_path));
        qn = qn > strip ? qn - strip : 0;
        if (qn != sn) die("sequence and quality have different lengths," + std::string(mb.ids.data() + mb.id_off[n], idn));
        memcpy(mb.qual + n * mb.stride, qp, qn);
        if (qn < mb.stride) memset(mb.qual + n * mb.stride + qn, 0, mb.stride - qn);
        n++;
    }
    return n;
}

// peprocess.cpp:1884-1908 (pair IDs) and :1207-1319 / seprocess.cpp:741-867 (quality system sanity)
void FilterRun::first_batch_checks(const HostBatch& b)
{
    if (pe_ && b.n > 0) {
        const MateBuf &a = b.m[0], &c = b.m[1];
        const size_t l1 = a.id_off[1] - a.id_off[0], l2 = c.id_off[1] - c.id_off[0];
        bool warn = l1 != l2;
        if (!warn) {
            int diff = 0;
            for (size_t i = 0; i < l1; i++) diff += a.ids[i] != c.ids[i];
            warn = diff > 1;
        }
        if (warn) std::cerr << "Warning:read ID in fq1 and fq2 seems not in pair, please check the input files if you are not sure" << std::endl;
    }
    // the reference runs this on the first batch (patchSize reads) a worker finishes
    const MateBuf& a = b.m[0];
    const size_t nchk = std::min<size_t>(b.n, (size_t)hp_.patch_size);
    int q1_exceed = 0, q1_normal = 0, q1_sum = 0, q2_exceed = 0, q2_normal = 0, q2_sum = 0;
    uint64_t bases = 0;
    const int other = hp_.quality_phred == 64 ? 33 : 64;
    for (size_t i = 0; i < nchk; i++) {
        const uint8_t* q = a.qual + i * a.stride;
        bases += a.len[i];
        for (int k = 0; k < a.len[i]; k++) {
            const int b1 = (int)q[k] - hp_.quality_phred, b2 = (int)q[k] - other;
            q1_sum += b1; q2_sum += b2;
            if (b1 >= 0 && b1 <= hp_.max_base_quality) q1_normal++; else if (b1 < -10 || b1 > hp_.max_base_quality + 10) q1_exceed++;
            if (b2 >= 0 && b2 <= hp_.max_base_quality) q2_normal++; else if (b2 < -10 || b2 > hp_.max_base_quality + 10) q2_exceed++;
        }
    }
    if (bases == 0) die("no data");
    const float r1 = (float)q1_normal / bases, r2 = (float)q2_normal / bases;
    const float m1 = (float)q1_sum / bases, m2 = (float)q2_sum / bases;
    int s1 = q1_exceed ? 0 : 1, s2 = q2_exceed ? 0 : 1;
    if (r1 > r2) s1 += 3; else if (r1 < r2) s2 += 3; else { s1 += 3; s2 += 3; }
    if (!(m1 < 10 || m1 > hp_.max_base_quality)) s1 += 2;
    if (!(m2 < 10 || m2 > hp_.max_base_quality)) s2 += 2;
    if (s1 - s2 < -3) die("base quality seems abnormal,please check the quality system parameter or fastq file");
    if (s1 - s2 < 0) std::cerr << "Warning:base quality seems abnormal,please check the quality system parameter or fastq file" << std::endl;
}

void FilterRun::ingest()
{
    // spaceNum: trailing whitespace of the very first line of fq1 (peprocess.cpp:2066-2076)
    {
        LineReader probe(hp_.fq1_path);
        const char* p; size_t n;
        if (probe.next(p, n)) {
            int sp = 0;
            while (n > 0 && isspace((unsigned char)p[n - 1])) { sp++; n--; }
            strip_gz_ = sp;
            if (probe.next(p, n)) stride_ = std::max<size_t>(16, round16(n));      // first read sets the initial row stride
        }
    }
    LineReader r1(hp_.fq1_path);
    LineReader* r2 = pe_ ? new LineReader(hp_.fq2_path) : nullptr;
    uint64_t seq_no = 0, first = 0;
    const size_t lanes = (size_t)snk_engine_lanes(engines_[0]);
    for (;;) {
        HostBatch* b;
        if (!free_q_.pop(b)) break;
        size_t n2 = 0;
        std::thread t2;
        const double tp0 = now_s();
        if (pe_) t2 = std::thread([&] { n2 = parse_mate(*r2, *b, 1, hp_.batch_reads); });
        const size_t n1 = parse_mate(r1, *b, 0, hp_.batch_reads);
        if (pe_) {
            t2.join();
            if (n1 != n2) die("reads number in fq1 and fq2 are different");
        }
        t_parse_ += now_s() - tp0;
        if (pe_) {
            const size_t s = std::max(b->m[0].stride, b->m[1].stride);
            for (int m = 0; m < 2; m++) if (b->m[m].stride != s) b->m[m].reserve(b->m[m].cap_reads, s);
        }
        stride_ = b->m[0].stride;
        if (n1 == 0) { free_q_.push(b); break; }
        b->n = (uint32_t)n1; b->seq_no = seq_no; b->first_index = first;
        b->gpu = (int)(seq_no % engines_.size());
        b->lane = (int)((seq_no / engines_.size()) % lanes);
        if (seq_no == 0) first_batch_checks(*b);
        first += n1; seq_no++;
        gpu_q_.push(b);
        if (n1 < hp_.batch_reads) break;
    }
    total_reads_ = first;
    n_batches_total_ = seq_no;
    delete r2;
    gpu_q_.close();
}

// submit on the batch's (gpu, lane); keep up to lanes*gpus batches in flight, retire the oldest
void FilterRun::gpu_stage()
{
    std::deque<HostBatch*> inflight;
    const size_t depth = inflight_depth();
    auto retire = [&] {
        HostBatch* d = inflight.front(); inflight.pop_front();
        const double t0 = now_s();
        engine_check(snk_engine_lane_sync(engines_[d->gpu], d->lane));
        t_gpu_wait_ += now_s() - t0;
        fmt_q_.push(d);
    };
    HostBatch* b;
    while (gpu_q_.pop(b)) {
        if (inflight.size() >= depth) retire();
        snk_batch b1 = {b->m[0].seq, b->m[0].qual, b->m[0].len, b->n, (uint32_t)b->m[0].stride};
        if (pe_) {
            snk_batch b2 = {b->m[1].seq, b->m[1].qual, b->m[1].len, b->n, (uint32_t)b->m[1].stride};
            engine_check(snk_filter_pe_async(engines_[b->gpu], b->lane, &b1, &b2, b->m[0].res, b->m[1].res, b->first_index));
        } else {
            engine_check(snk_filter_se_async(engines_[b->gpu], b->lane, &b1, b->m[0].res, b->first_index));
        }
        inflight.push_back(b);
    }
    while (!inflight.empty()) retire();
    fmt_q_.close();
}

void FilterRun::encode(std::string& text)
{
    if (!hp_.clean_gz || text.empty()) return;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, 2, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) die("zlib deflateInit2 failed");
    std::string out;
    out.resize(deflateBound(&zs, (uLong)text.size()) + 32);
    zs.next_in = (Bytef*)text.data(); zs.avail_in = (uInt)text.size();
    zs.next_out = (Bytef*)&out[0]; zs.avail_out = (uInt)out.size();
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) die("zlib deflate failed");
    out.resize(zs.total_out);
    deflateEnd(&zs);
    text.swap(out);
}

// Formats the surviving records of one mate. Output is a list of pieces so that the ordered writer
// can reproduce the reference's emission order (see FilterRun::writer).
void FilterRun::format_mate(HostBatch& b, int mate)
{
    MateBuf& mb = b.m[mate];
    mb.out.clear();
    std::string cur;
    int cur_kind = 0;
    auto flush_piece = [&](int kind) {
        if (!cur.empty()) { encode(cur); mb.out.push_back({cur_kind, std::move(cur)}); cur.clear(); }
        cur_kind = kind;
    };
    const int shift = hp_.out_quality_phred - hp_.quality_phred;
    const bool fasta = hp_.output_file_type == "fasta";
    cur.reserve((size_t)b.n * (mb.stride * 2 + 64) / 1);
    for (uint32_t i = 0; i < b.n; i++) {
        const uint64_t gi = b.first_index + i;
        int kind = 0;
        if (reorder_) {
            const uint64_t in_cyc = gi % cyc_;
            if (gi >= cyc_ && in_cyc == insert_off_) { flush_piece(cur_kind); mb.out.push_back({2, std::string()}); }
            if (in_cyc >= cyc_ - defer_len_) kind = 1;     // if the input ends inside this range it is flushed at EOF: same order
        }
        if (kind != cur_kind) flush_piece(kind);
        const snk_read_result& r = mb.res[i];
        if (r.category != SNK_KEEP) continue;
        const char* id = mb.ids.data() + mb.id_off[i];
        size_t idn = mb.id_off[i + 1] - mb.id_off[i];
        std::string idbuf;
        if (hp_.index_remove) {                      // read_filter.cpp:357-382
            if (hp_.seq_type == "0") {
                bool cp = true;
                for (size_t k = 0; k < idn; k++) {
                    if (id[k] == '#') cp = false;
                    if (cp) idbuf += id[k];
                    else if (id[k] == '/') { cp = true; idbuf += id[k]; }
                }
            } else {
                idbuf.assign(id, idn);
                const size_t c = idbuf.find_last_of(':');
                idbuf = idbuf.substr(0, c);           // npos -> whole string, as substr(0, npos)
            }
            id = idbuf.data(); idn = idbuf.size();
        }
        const uint8_t* s = mb.seq + (size_t)i * mb.stride + r.head_cut;
        const uint8_t* q = mb.qual + (size_t)i * mb.stride + r.head_cut;
        if (fasta) {
            std::string t(id, idn);
            const size_t at = t.find('@');
            if (at != std::string::npos) t[at] = '>';
            cur += t;
            if (hp_.pe_info) cur += mate ? "/2" : "/1";
            cur += '\n';
            cur.append((const char*)s, r.clean_len);
            cur += '\n';
            continue;
        }
        cur.append(id, idn);
        if (hp_.pe_info) cur += mate ? "/2" : "/1";
        cur += '\n';
        cur.append((const char*)s, r.clean_len);
        cur += "\n+\n";
        if (shift == 0) cur.append((const char*)q, r.clean_len);
        else for (int k = 0; k < r.clean_len; k++) cur += (char)((int)q[k] + shift);
        cur += '\n';
    }
    flush_piece(0);
}

void FilterRun::format_worker()
{
    HostBatch* b;
    while (fmt_q_.pop(b)) {
        const double t0 = now_s();
        for (int m = 0; m < mates_; m++) format_mate(*b, m);
        t_format_us_ += (uint64_t)((now_s() - t0) * 1e6);
        { std::lock_guard<std::mutex> g(done_mu_); done_[b->seq_no] = b; }
        done_cv_.notify_all();
    }
}

// Ordered writer. Emission order of the reference (peprocess.cpp:2141,2248,2957-2990): worker i owns
// the blocks b with b % T == i and appends each batch of patchSize reads to a temp file named by
// (worker, cycle); the files are concatenated cycle-major, worker-minor. The cycle label of a full
// batch comes from a line counter that, for plain-text PE input, has already moved past the batch,
// so the last patchSize reads before every cycle boundary are labelled with the next cycle and come
// out right before the last worker's block of that next cycle (or at the very end). Pieces of kind 1
// hold such deferred records, markers of kind 2 are those insertion points. Everything else (and
// all .gz-input and all SE runs) is input order.
void FilterRun::writer()
{
    FILE* out[2] = {nullptr, nullptr};
    const std::string names[2] = {hp_.output_dir + "/" + hp_.clean_fq1, hp_.output_dir + "/" + hp_.clean_fq2};
    for (int m = 0; m < mates_; m++) {
        out[m] = fopen(names[m].c_str(), "wb");
        if (!out[m]) die("cannot write to the file," + names[m]);
        setvbuf(out[m], nullptr, _IOFBF, 1 << 22);
    }
    uint64_t next = 0;
    for (;;) {
        HostBatch* b = nullptr;
        {
            std::unique_lock<std::mutex> g(done_mu_);
            done_cv_.wait(g, [&] { return done_.count(next) || fmt_finished_; });
            auto it = done_.find(next);
            if (it == done_.end()) break;
            b = it->second; done_.erase(it);
        }
        const double tw0 = now_s();
        for (int m = 0; m < mates_; m++) {
            for (Piece& p : b->m[m].out) {
                if (p.kind == 0) fwrite(p.bytes.data(), 1, p.bytes.size(), out[m]);
                else if (p.kind == 1) pending_deferred_[m] += p.bytes;
                else { fwrite(pending_deferred_[m].data(), 1, pending_deferred_[m].size(), out[m]); pending_deferred_[m].clear(); }
            }
            b->m[m].out.clear();
        }
        t_write_ += now_s() - tw0;
        if (b->seq_no % 4 == 0) log_line(local_time() + " processed_reads:\t" + std::to_string(b->first_index + b->n));
        next++;
        free_q_.push(b);
    }
    // End of input. The reference's final concat pass (peprocess.cpp:2957-2966) walks the workers of
    // the last cycle in order and stops at the first one without a temp file for it; the deferred
    // batch sits in the LAST worker's file, so it is silently dropped (while still counted in the
    // clean statistics) whenever the input ends at or before worker T-2's block of the final cycle.
    // Reproduced for byte parity; SNK_KEEP_DEFERRED=1 writes those records instead of losing them.
    bool drop = false;
    if (reorder_ && total_reads_ >= cyc_ && !getenv("SNK_KEEP_DEFERRED")) {
        const uint64_t into_last = total_reads_ - (total_reads_ / cyc_) * cyc_;
        drop = into_last <= (uint64_t)ep_.slot_block * (uint64_t)(ep_.n_slots - 2);
    }
    for (int m = 0; m < mates_; m++) {
        if (!drop) fwrite(pending_deferred_[m].data(), 1, pending_deferred_[m].size(), out[m]);
        if (fclose(out[m]) != 0) die("cannot write to the file," + names[m]);
    }
}

void FilterRun::process()
{
    const double t_begin = now_s();
    mkdir_p(hp_.output_dir);
    log_.open(hp_.log.c_str());
    if (!log_) die("cannot open such file," + hp_.log);
    log_line(local_time() + "\tAnalysis start!");
    if (snk_params_check(&ep_)) die(snk_last_error());
    for (int g = 0; g < hp_.n_gpus; g++) {
        snk_engine* e = nullptr;
        engine_check(snk_engine_create(&ep_, g, &e));
        engines_.push_back(e);
    }
    // emission-order quirk applies to plain-text PE input with more than one worker
    cyc_ = (uint64_t)ep_.slot_block * (uint64_t)ep_.n_slots;
    defer_len_ = (uint64_t)hp_.patch_size;
    insert_off_ = (uint64_t)ep_.slot_block * (uint64_t)(ep_.n_slots - 1);
    reorder_ = pe_ && !hp_.input_gz && ep_.n_slots > 1;
    batches_.resize(inflight_depth() + 3);       // in flight on the GPUs + being parsed + being formatted/written
    for (auto& b : batches_) free_q_.push(&b);

    t_setup_ = now_s() - t_begin;
    const int nworkers = std::max(2, hp_.threads);
    std::thread t_ingest([&] { ingest(); });
    std::thread t_gpu([&] { gpu_stage(); });
    std::vector<std::thread> workers;
    for (int i = 0; i < nworkers; i++) workers.emplace_back([&] { format_worker(); });
    std::thread t_writer([&] { writer(); });
    t_ingest.join();
    t_gpu.join();
    for (auto& w : workers) w.join();
    { std::lock_guard<std::mutex> g(done_mu_); fmt_finished_ = true; }
    done_cv_.notify_all();
    t_writer.join();
    free_q_.close();

    // ---- statistics: per-GPU tables -> one table (counters add, LAST_KEY words take the max)
    const size_t words = (size_t)ep_.n_slots * SNK_SLOT_WORDS;
    std::vector<uint64_t> total(words, 0), part(words);
    std::vector<size_t> key_words;
    for (int s = 0; s < ep_.n_slots; s++)
        for (int f = 0; f < SNK_FILE_COUNT; f++)
            key_words.push_back((size_t)s * SNK_SLOT_WORDS + SNK_SLOT_FILE_OFF(f) + SNK_FILE_GS_OFF + SNK_GS_LAST_KEY);
    for (snk_engine* e : engines_) {
        uint32_t flags = 0; uint64_t bad = 0;
        engine_check(snk_engine_error_flags(e, &flags, &bad));
        if (flags & 1) die("unrecognized sequence, read number " + std::to_string(bad + 1));
        if (flags & 2) die("base quality is out of range,please check the quality system parameter or fastq file, read number " + std::to_string(bad + 1));
        if (flags & 4) die("low quality base ratio stat error, read number " + std::to_string(bad + 1));
        engine_check(snk_engine_stats(e, part.data()));
        std::vector<uint64_t> keys;
        for (size_t k : key_words) { keys.push_back(std::max(total[k], part[k])); part[k] = 0; total[k] = 0; }
        for (size_t i = 0; i < words; i++) total[i] += part[i];
        for (size_t j = 0; j < key_words.size(); j++) total[key_words[j]] = keys[j];
    }
    if (pe_) { if (snk_report_write_pe(&ep_, total.data(), hp_.output_dir.c_str())) die(snk_last_error()); }
    else { if (snk_report_write_se(&ep_, total.data(), hp_.output_dir.c_str())) die(snk_last_error()); }
    for (snk_engine* e : engines_) snk_engine_destroy(e);
    engines_.clear();
    for (auto& b : batches_) for (auto& m : b.m) m.release();
    batches_.clear();
    {
        char buf[512];
        snprintf(buf, sizeof buf, "stage seconds: setup %.2f, parse(busy) %.2f, gpu-wait %.2f, format(sum over %d workers) %.2f, write %.2f, total %.2f; reads %llu",
                 t_setup_, t_parse_, t_gpu_wait_, std::max(2, hp_.threads), t_format_us_.load() * 1e-6, t_write_, now_s() - t_begin,
                 (unsigned long long)total_reads_);
        log_line(buf);
    }
    log_line(local_time() + "\tAnalysis accomplished!");
    log_.close();
}

peProcess::peProcess(const HostParams& hp) : run_(new FilterRun(hp, true)) {}
peProcess::~peProcess() { delete run_; }
void peProcess::process() { run_->process(); }
seProcess::seProcess(const HostParams& hp) : run_(new FilterRun(hp, false)) {}
seProcess::~seProcess() { delete run_; }
void seProcess::process() { run_->process(); }

} // namespace snk

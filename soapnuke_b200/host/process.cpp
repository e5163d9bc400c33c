// process.cpp — the pinned-buffer batching driver behind peProcess::process / seProcess::process.
// See process.h for the stage diagram. Reference behaviour reproduced here (file:line):
//   line handling of sub_thread           peprocess.cpp:2066-2076, 2090-2131 (.gz: strip the first line's
//                                         trailing-whitespace count from every line), :2198-2239 (plain: strip 1)
//   first-batch pair-ID / Phred checks    peprocess.cpp:1884-1908, 1207-1319; seprocess.cpp:741-867
//   record formatting                     on the device (csrc/text_kernels.cuh): peprocess.cpp:3383-3433, :1617-1629,
//                                         read_filter.cpp:357-382
//   emission order of the clean records   peprocess.cpp:2141,2248,2957-2990 (see Writer::route)
//   gzip level 2 members                  peprocess.cpp:1803-1810
#include "process.h"
#include "host_common.h"
#include "gz_members.h"
#include "fast_deflate.h"
#include "../csrc/text_core.cuh"      // id_transform (host-compilable header; the record formatting of the trim files)
#include <zlib.h>
#include <nvtx3/nvToolsExt.h>     // header-only; ranges cost nothing unless a timeline tool is attached
#include <sys/stat.h>
#include <sys/mman.h>
#include <sys/vfs.h>
#include <fcntl.h>
#include <unistd.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <algorithm>
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
#include <memory>
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

struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

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

size_t count_newlines(const char* p, size_t n);

// ------------------------------------------------------------------ raw byte source (plain via read(2), .gz via zlib)
class ByteSource {
public:
    // gz_threads > 0: multi-member .gz files are inflated member by member on that many host threads (gz_members.h)
    ByteSource(const std::string& path, bool gz, int gz_threads = 0) : path_(path), gz_(gz)
    {
        if (gz_ && gz_threads > 0 && !getenv("SNK_GZ_SERIAL")) members_.reset(GzMemberReader::open(path, gz_threads));
        if (members_) return;
        if (gz_) {
            f_ = gzopen(path.c_str(), "rb");
            if (!f_) die("cannot open the file," + path);
            gzbuffer(f_, 1 << 22);
        } else {
            fd_ = open(path.c_str(), O_RDONLY);
            if (fd_ < 0) die("cannot open the file," + path);
            seekable_ = lseek(fd_, 0, SEEK_CUR) != (off_t)-1;
#ifdef F_SETPIPE_SZ
            if (!seekable_) fcntl(fd_, F_SETPIPE_SZ, 1 << 20);     // a FIFO / pipe: 1 MiB of buffering instead of 64 KiB
#endif
        }
    }
    ~ByteSource() { stop_prefetch(); if (pf_buf_) free(pf_buf_); if (f_) gzclose(f_); if (fd_ >= 0) close(fd_); }
    // Read-ahead for the seconds in which nothing else can happen: creating the CUDA contexts takes 0.3 - 2 s and the
    // pinned batch buffers cannot be allocated before it is done. A plain, seekable input is meanwhile copied out of the
    // page cache into ordinary memory (up to max_bytes); read() then serves those offsets with memcpy. stop_prefetch()
    // is called when the first batch buffer becomes available - from there on reading ahead would only copy twice.
    void start_prefetch(size_t max_bytes, int threads)
    {
        if (gz_ || members_ || !seekable_ || fd_ < 0 || max_bytes == 0) return;
        struct stat st;
        if (fstat(fd_, &st) != 0 || st.st_size <= 0) return;
        pf_cap_ = std::min<size_t>(max_bytes, (size_t)st.st_size);
        pf_buf_ = (char*)malloc(pf_cap_);
        if (!pf_buf_) { pf_cap_ = 0; return; }
        pf_thread_ = std::thread([this, threads] {
            constexpr size_t kStep = 32u << 20;
            const int np = std::max(1, threads);
            size_t done = 0;
            while (done < pf_cap_ && !pf_stop_.load(std::memory_order_acquire)) {
                const size_t n = std::min(kStep, pf_cap_ - done);
                const size_t part = (n / np + 4095) & ~(size_t)4095;
                std::vector<std::thread> th;
                std::atomic<bool> short_read{false};
                for (int i = 0; i < np; i++) {
                    const size_t lo = std::min(n, part * i), hi = std::min(n, part * (i + 1));
                    if (lo == hi) continue;
                    th.emplace_back([&, lo, hi] {
                        size_t got = 0;
                        while (got < hi - lo) {
                            const ssize_t g = ::pread(fd_, pf_buf_ + done + lo + got, hi - lo - got, (off_t)(done + lo + got));
                            if (g <= 0) { short_read = true; break; }
                            got += (size_t)g;
                        }
                    });
                }
                for (auto& t : th) t.join();
                if (short_read) break;                      // the file changed under us: the rest is read the normal way
                done += n;
                pf_done_.store(done, std::memory_order_release);
            }
        });
    }
    void stop_prefetch()
    {
        pf_stop_.store(true, std::memory_order_release);
        if (pf_thread_.joinable()) pf_thread_.join();
    }
    size_t prefetched() const { return pf_done_.load(std::memory_order_acquire); }
    // up to n bytes into dst; 0 at EOF. Plain files: large requests are split over a few threads (the copy out of the
    // page cache is what limits a single reader); with `parts` every thread also counts the newlines of its share, so
    // that the caller's search for a record boundary only has to look into one share.
    struct Part { size_t off, len, newlines; };
    size_t read(char* dst, size_t n, std::vector<Part>* parts = nullptr)
    {
        if (parts) parts->clear();
        if (members_) {
            const size_t got = members_->read(dst, n);
            if (got == GzMemberReader::kError) die("cannot read the file," + path_);
            return got;
        }
        if (gz_) {
            int got = gzread(f_, dst, (unsigned)std::min<size_t>(n, 1u << 30));
            if (got < 0) die("cannot read the file," + path_);
            return (size_t)got;
        }
        if (!seekable_) {                       // a pipe: plain sequential reads
            const ssize_t got = ::read(fd_, dst, n);
            if (got < 0) die("cannot read the file," + path_);
            return (size_t)got;
        }
        auto pread_all = [&](char* d, size_t want, off_t at) -> size_t {
            if ((size_t)at + want <= pf_done_.load(std::memory_order_acquire)) { memcpy(d, pf_buf_ + at, want); return want; }   // read ahead earlier
            size_t done = 0;
            while (done < want) {
                const ssize_t got = ::pread(fd_, d + done, want - done, at + (off_t)done);
                if (got < 0) die("cannot read the file," + path_);
                if (got == 0) break;
                done += (size_t)got;
            }
            return done;
        };
        size_t total = 0;
        if (n < (4u << 20)) total = pread_all(dst, n, off_);
        else {
            const int np = (int)std::min<size_t>((size_t)read_threads_, n >> 20);
            const size_t part = (n / np + 4095) & ~(size_t)4095;
            std::vector<size_t> got(np, 0), nl(np, 0);
            std::vector<std::thread> th(np);
            auto work = [&](int i) {
                const size_t lo = std::min(n, part * i), hi = std::min(n, part * (i + 1));
                got[i] = pread_all(dst + lo, hi - lo, off_ + (off_t)lo);
                if (parts) nl[i] = count_newlines(dst + lo, got[i]);
            };
            for (int i = 1; i < np; i++) th[i] = std::thread(work, i);
            work(0);
            for (int i = 1; i < np; i++) th[i].join();
            for (int i = 0; i < np; i++) {
                const size_t lo = std::min(n, part * i), hi = std::min(n, part * (i + 1));
                if (parts && got[i]) parts->push_back({lo, got[i], nl[i]});
                total += got[i];
                if (got[i] < hi - lo) break;      // end of file inside this part
            }
        }
        off_ += (off_t)total;
        return total;
    }
    void set_read_threads(int n) { read_threads_ = std::max(1, n); }
    // bytes carried over from the previous batch (text after its last complete record)
    std::vector<char> carry;
    bool eof = false;
    bool parallel_gz() const { return (bool)members_; }
    GzMemberReader::Counters gz_counters() const { return members_ ? members_->counters() : GzMemberReader::Counters(); }
private:
    std::unique_ptr<GzMemberReader> members_;
    std::string path_;
    bool gz_;
    gzFile f_ = nullptr;
    int fd_ = -1;
    off_t off_ = 0;
    bool seekable_ = true;
    int read_threads_ = 4;
    char* pf_buf_ = nullptr;
    size_t pf_cap_ = 0;
    std::atomic<size_t> pf_done_{0};
    std::atomic<bool> pf_stop_{false};
    std::thread pf_thread_;
};

// ------------------------------------------------------------------ newline search
// Offset just behind the `need`-th '\n' of p[0,n), or n when there are fewer; *count = newlines seen
// (stops counting at `need`).
#if defined(__x86_64__)
__attribute__((target("avx2"))) size_t nth_newline_avx2(const char* p, size_t n, size_t need, size_t* count)
{
    size_t i = 0, c = 0;
    const __m256i nl = _mm256_set1_epi8('\n');
    while (i + 32 <= n) {
        const uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)(p + i)), nl));
        const size_t k = (size_t)__builtin_popcount(m);
        if (c + k >= need) {
            uint32_t mm = m;
            for (size_t skip = need - c - 1; skip > 0; skip--) mm &= mm - 1;
            *count = need;
            return i + (size_t)__builtin_ctz(mm) + 1;
        }
        c += k;
        i += 32;
    }
    for (; i < n; i++)
        if (p[i] == '\n' && ++c == need) { *count = c; return i + 1; }
    *count = c;
    return n;
}
#endif
#if defined(__x86_64__)
__attribute__((target("avx2"))) size_t count_newlines_avx2(const char* p, size_t n)
{
    size_t i = 0, c = 0;
    const __m256i nl = _mm256_set1_epi8('\n');
    for (; i + 32 <= n; i += 32)
        c += (size_t)__builtin_popcount((uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)(p + i)), nl)));
    for (; i < n; i++) c += p[i] == '\n';
    return c;
}
#endif
size_t count_newlines(const char* p, size_t n)
{
#if defined(__x86_64__)
    static const bool has_avx2 = __builtin_cpu_supports("avx2");
    if (has_avx2) return count_newlines_avx2(p, n);
#endif
    size_t c = 0;
    for (size_t i = 0; i < n; i++) c += p[i] == '\n';
    return c;
}
size_t nth_newline(const char* p, size_t n, size_t need, size_t* count)
{
    if (need == 0) { *count = 0; return 0; }
#if defined(__x86_64__)
    static const bool has_avx2 = __builtin_cpu_supports("avx2");
    if (has_avx2) return nth_newline_avx2(p, n, need, count);
#endif
    size_t i = 0, c = 0;
    while (i < n) {
        const char* q = (const char*)memchr(p + i, '\n', n - i);
        if (!q) break;
        i = (size_t)(q - p) + 1;
        if (++c == need) { *count = c; return i; }
    }
    *count = c;
    return n;
}

// ------------------------------------------------------------------ one batch travelling through the stages
struct PinnedBuf {
    char* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes, size_t keep)          // contents [0, keep) survive
    {
        if (bytes <= cap) return;
        size_t want = bytes + bytes / 4 + (1u << 20);
        char* np = nullptr;
        engine_check(snk_host_alloc((void**)&np, want));
        if (p) { if (keep) memcpy(np, p, keep); snk_host_free(p); }
        p = np; cap = want;
    }
    void release() { if (p) snk_host_free(p); p = nullptr; cap = 0; }
};
// a run of output bytes: kind 0 = in place, 1 = deferred (see FilterRun::writer), 2 = "emit the deferred bytes now"
// Trim pieces (trimFq1/2) carry a record range instead of bytes until a worker has formatted them.
struct Piece { int kind; const char* p; size_t len; std::string gz; uint32_t r0 = 0, r1 = 0; uint32_t kept = 0; /* records in a deferred clean piece */ };
struct HostBatch {
    uint64_t seq_no = 0, first_index = 0;
    uint32_t n = 0;
    int gpu = 0, lane = 0;
    PinnedBuf in[2], out[2], off[2];      // raw text, clean text, rec_off (uint32 [n+1])
    size_t in_bytes[2] = {0, 0};
    snk_text_meta meta;
    PinnedBuf res[2];                     // per-read results (only fetched when the trim files are written)
    std::vector<uint32_t> rec_start[2];   // byte offset of every record in the raw text, n+1 entries (trim files only)
    std::vector<Piece> pieces[2], tpieces[2];   // clean / trim output runs
    std::atomic<int> tasks{0};            // deflate tasks still running
    void release() { for (int m = 0; m < 2; m++) { in[m].release(); out[m].release(); off[m].release(); res[m].release(); } }
};
struct GzTask { HostBatch* b; int mate; size_t piece; bool trim; };

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
    snk_text_format fmt_;
    std::vector<snk_engine*> engines_;
    std::ofstream log_;
    std::mutex log_mu_;
    int strip_gz_ = 1;                 // spaceNum of the first line (peprocess.cpp:2066-2076)
    size_t stride_ = 0;                // current row stride (grows when a longer read shows up)
    uint64_t total_reads_ = 0;
    std::string pending_deferred_[4];  // deferred output (already encoded) waiting for its insertion point; [2..3] = trim files
    bool trim_ = false;                // trimFq1/2: every record after trimming, before the discard decision
    uint64_t deferred_records_ = 0;    // kept records (pairs) waiting in pending_deferred_[0]

    std::deque<HostBatch> batches_;
    Channel<HostBatch*> free_q_, gpu_q_;
    Channel<GzTask> gz_q_;
    std::mutex done_mu_;
    std::condition_variable done_cv_;
    std::map<uint64_t, HostBatch*> done_;       // finished batches by sequence number
    bool all_submitted_ = false;
    uint64_t n_batches_total_ = 0;

    // stages
    void ingest();
    size_t fill_mate(ByteSource& src, HostBatch& b, int mate, size_t max_reads);
    void first_batch_checks(const HostBatch& b);
    void gpu_stage();
    void make_pieces(HostBatch& b);
    template <class F> void walk_runs(const HostBatch& b, F&& emit);
    void index_records(HostBatch& b, int mate);
    void format_trim(const HostBatch& b, int mate, uint32_t r0, uint32_t r1, std::string& out) const;
    void finish_batch(HostBatch* b);
    void gz_worker();
    void writer();
    void encode(const char* p, size_t n, std::string& out);
    size_t inflight_depth() const { return engines_.size() * (size_t)snk_engine_lanes(engines_[0]); }
    void log_line(const std::string& s) { std::lock_guard<std::mutex> g(log_mu_); log_ << s << std::endl; }

    // output routing (reference emission order)
    uint64_t cyc_ = 0, defer_len_ = 0, insert_off_ = 0;
    bool reorder_ = false;
    // busy seconds per stage (log only)
    double t_read_ = 0, t_gpu_wait_ = 0, t_setup_ = 0;
    double tl_prefetched_ = 0;           // GB read ahead while the engines came up
    double t0_ = 0, tl_first_batch_ = 0, tl_ingest_done_ = 0, tl_gpu_done_ = 0, tl_writer_done_ = 0, tl_stats_done_ = 0;   // timeline marks (log only)
    std::atomic<uint64_t> t_gz_us_{0}, t_write_us_{0};
    size_t avg_rec_bytes_[2] = {0, 0};
};

// ---- the raw text of up to max_reads whole records of one mate, straight into the pinned buffer
size_t FilterRun::fill_mate(ByteSource& src, HostBatch& b, int mate, size_t max_reads)
{
    PinnedBuf& buf = b.in[mate];
    const size_t want_lines = 4 * max_reads;
    const size_t guess = avg_rec_bytes_[mate] ? avg_rec_bytes_[mate] : (2 * (stride_ ? stride_ : 160) + 64);
    size_t have = src.carry.size();
    buf.reserve(std::max(have, max_reads * guess) + (1u << 20), 0);
    memcpy(buf.p, src.carry.data(), have);
    src.carry.clear();
    size_t scanned = 0, lines = 0, end = 0;
    bool complete = false;
    std::vector<ByteSource::Part> parts;
    for (;;) {
        if (scanned < have) {
            size_t c = 0;
            const size_t at = nth_newline(buf.p + scanned, have - scanned, want_lines - lines, &c);
            lines += c;
            if (lines == want_lines) { end = scanned + at; complete = true; break; }
            scanned = have;
        }
        if (src.eof) break;
        // read about what is still missing (plus a little), never less than 1 MiB
        size_t missing = (want_lines - lines) / 4 * guess + (256u << 10);
        if (missing < (1u << 20)) missing = 1u << 20;
        buf.reserve(have + missing + 64, have);
        const size_t got = src.read(buf.p + have, missing, &parts);
        if (got == 0) src.eof = true;
        // shares whose newlines were counted by the reading threads: skip the ones that cannot hold the boundary
        for (const ByteSource::Part& pt : parts) {
            if (pt.off != scanned - have || lines + pt.newlines >= want_lines) break;
            lines += pt.newlines; scanned += pt.len;
        }
        have += got;
        if (have > 0xE0000000ull) die("batch text exceeds 3.5 GiB: lower SNK_BATCH_READS");
    }
    if (!complete) {                   // end of input: everything that is left
        end = have;
        if (have > 0 && buf.p[have - 1] != '\n') lines++;      // last line without '\n'
        if (lines % 4 != 0) die("input fastq is truncated," + (mate ? hp_.fq2_path : hp_.fq1_path));
    } else {
        src.carry.assign(buf.p + end, buf.p + have);
    }
    b.in_bytes[mate] = end;
    const size_t n = lines / 4;
    if (n) avg_rec_bytes_[mate] = end / n + 1;
    return n;
}

// peprocess.cpp:1884-1908 (pair IDs) and :1207-1319 / seprocess.cpp:741-867 (quality system sanity),
// on the first patchSize records of the run, parsed here from the raw text
void FilterRun::first_batch_checks(const HostBatch& b)
{
    const size_t strip = (size_t)fmt_.strip;
    auto line_at = [&](int m, size_t& pos, const char*& p, size_t& n) {
        const char* base = b.in[m].p;
        const size_t end = b.in_bytes[m];
        const char* nl = pos < end ? (const char*)memchr(base + pos, '\n', end - pos) : nullptr;
        const size_t raw = nl ? (size_t)(nl - (base + pos)) + 1 : end - pos;
        p = base + pos; n = raw > strip ? raw - strip : 0;
        pos += raw;
    };
    if (pe_ && b.n > 0) {
        size_t p1 = 0, p2 = 0, l1, l2; const char *a, *c;
        line_at(0, p1, a, l1); line_at(1, p2, c, l2);
        bool warn = l1 != l2;
        if (!warn) {
            int diff = 0;
            for (size_t i = 0; i < l1; i++) diff += a[i] != c[i];
            warn = diff > 1;
        }
        if (warn) std::cerr << "Warning:read ID in fq1 and fq2 seems not in pair, please check the input files if you are not sure" << std::endl;
    }
    // the reference runs this on the first batch (patchSize reads) a worker finishes
    const size_t nchk = std::min<size_t>(b.n, (size_t)hp_.patch_size);
    int q1_exceed = 0, q1_normal = 0, q1_sum = 0, q2_exceed = 0, q2_normal = 0, q2_sum = 0;
    uint64_t bases = 0;
    const int other = hp_.quality_phred == 64 ? 33 : 64;
    size_t pos = 0;
    for (size_t i = 0; i < nchk; i++) {
        const char *idp, *sp, *pp, *qp; size_t idn, sn, pn, qn;
        line_at(0, pos, idp, idn); line_at(0, pos, sp, sn); line_at(0, pos, pp, pn); line_at(0, pos, qp, qn);
        const size_t len = std::min(sn, qn);
        bases += len;
        for (size_t k = 0; k < len; k++) {
            const int b1 = (int)(uint8_t)qp[k] - hp_.quality_phred, b2 = (int)(uint8_t)qp[k] - other;
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
    // spaceNum: trailing whitespace of the very first line of fq1 (peprocess.cpp:2066-2076); the
    // reference probes it with gzopen/gzgets, which also reads plain files
    // .gz input: the host threads (-T) are shared by the mates' member decoders
    const int gz_threads = hp_.input_gz ? std::max(1, hp_.threads / mates_) : 0;
    struct stat st1;
    const bool fifo1 = stat(hp_.fq1_path.c_str(), &st1) == 0 && !S_ISREG(st1.st_mode);
    std::unique_ptr<ByteSource> r1_early;
    {
        // a pipe / FIFO cannot be opened twice: probe through the real reader and hand the bytes back to it as carry
        std::vector<char> head(1 << 16);
        size_t got = 0;
        if (fifo1) {
            r1_early.reset(new ByteSource(hp_.fq1_path, hp_.input_gz, gz_threads));
            while (got < head.size()) { const size_t g = r1_early->read(head.data() + got, head.size() - got); if (!g) break; got += g; }
            r1_early->carry.assign(head.data(), head.data() + got);
        } else {
            ByteSource probe(hp_.fq1_path, true);
            got = probe.read(head.data(), head.size());
        }
        const char* nl = (const char*)memchr(head.data(), '\n', got);
        size_t n = nl ? (size_t)(nl - head.data()) + 1 : got;
        int sp = 0;
        while (n > 0 && isspace((unsigned char)head[n - 1])) { sp++; n--; }
        strip_gz_ = sp;
        if (nl) {                                                  // first read sets the initial row stride
            const char* nl2 = (const char*)memchr(nl + 1, '\n', got - (size_t)(nl + 1 - head.data()));
            if (nl2) stride_ = std::max<size_t>(16, round16((size_t)(nl2 - nl)));
        }
    }
    if (!stride_) stride_ = 160;
    fmt_.strip = hp_.input_gz ? strip_gz_ : 1;                     // plain: erase(size()-1) (peprocess.cpp:2206)
    std::unique_ptr<ByteSource> r1_own(r1_early ? r1_early.release() : new ByteSource(hp_.fq1_path, hp_.input_gz, gz_threads));
    ByteSource& r1 = *r1_own;
    ByteSource* r2 = pe_ ? new ByteSource(hp_.fq2_path, hp_.input_gz, gz_threads) : nullptr;
    {   // plain files: threads per mate that copy out of the page cache (and count newlines) in parallel
        int rt = std::max(2, std::min(8, hp_.threads / mates_));
        if (const char* e = getenv("SNK_READ_THREADS")) rt = std::max(1, atoi(e));
        r1.set_read_threads(rt);
        if (r2) r2->set_read_threads(rt);
    }
    // the engines (CUDA contexts) are still being created by process(): read ahead until the first batch buffer arrives
    {
        size_t pf_mb = 2048;
        if (const char* e = getenv("SNK_PREFETCH_MB")) pf_mb = (size_t)std::max(0, atoi(e));
        const int rt = std::max(2, std::min(8, hp_.threads / mates_));
        r1.start_prefetch(pf_mb << 20, rt);
        if (r2) r2->start_prefetch(pf_mb << 20, rt);
    }
    uint64_t seq_no = 0, first = 0;
    const size_t lanes = (size_t)snk_engine_lanes(nullptr);
    bool prefetching = true;
    for (;;) {
        HostBatch* b;
        if (!free_q_.pop(b)) break;
        if (prefetching) {
            r1.stop_prefetch();
            if (r2) r2->stop_prefetch();
            prefetching = false;
            tl_prefetched_ = (double)(r1.prefetched() + (r2 ? r2->prefetched() : 0)) / 1e9;
        }
        size_t n2 = 0;
        std::thread t2;
        NvtxRange nvtx("snk:ingest_batch");
        const double tp0 = now_s();
        if (pe_) t2 = std::thread([&] { n2 = fill_mate(*r2, *b, 1, hp_.batch_reads); });
        const size_t n1 = fill_mate(r1, *b, 0, hp_.batch_reads);
        if (pe_) {
            t2.join();
            if (n1 != n2) die("reads number in fq1 and fq2 are different");
        }
        t_read_ += now_s() - tp0;
        if (n1 == 0) { free_q_.push(b); break; }
        b->n = (uint32_t)n1; b->seq_no = seq_no; b->first_index = first;
        b->gpu = (int)(seq_no % engines_.size());
        b->lane = (int)((seq_no / engines_.size()) % lanes);
        if (seq_no == 0) { first_batch_checks(*b); tl_first_batch_ = now_s() - t0_; }
        first += n1; seq_no++;
        gpu_q_.push(b);
        if (n1 < hp_.batch_reads) break;
    }
    total_reads_ = first;
    n_batches_total_ = seq_no;
    tl_ingest_done_ = now_s() - t0_;
    if (r1.parallel_gz()) {
        const GzMemberReader::Counters c1 = r1.gz_counters(), c2 = r2 ? r2->gz_counters() : GzMemberReader::Counters();
        log_line("gzip input: " + std::to_string(c1.members + c2.members) + " members inflated on " + std::to_string(gz_threads * mates_) +
                 " host threads, " + std::to_string(c1.cancelled + c2.cancelled) + " false member candidates dropped");
    }
    delete r2;
    gpu_q_.close();
}

// submit on the batch's (gpu, lane); keep up to lanes*gpus batches in flight, retire the oldest
void FilterRun::gpu_stage()
{
    std::deque<HostBatch*> inflight;
    const size_t depth = inflight_depth();
    auto submit = [&](HostBatch* b) {
        NvtxRange nvtx("snk:gpu_submit");
        if (pe_) engine_check(snk_filter_pe_text_async(engines_[b->gpu], b->lane, b->in[0].p, b->in_bytes[0], b->in[1].p, b->in_bytes[1],
                                                       b->n, (uint32_t)stride_, &fmt_, b->first_index));
        else engine_check(snk_filter_se_text_async(engines_[b->gpu], b->lane, b->in[0].p, b->in_bytes[0], b->n, (uint32_t)stride_, &fmt_,
                                                   b->first_index));
    };
    auto retire = [&] {
        HostBatch* d = inflight.front(); inflight.pop_front();
        NvtxRange nvtx("snk:gpu_retire");
        const double t0 = now_s();
        for (;;) {
            engine_check(snk_text_meta_sync(engines_[d->gpu], d->lane, &d->meta));
            if (d->meta.flags & SNK_TEXT_TOO_LONG) die("read longer than 1000 bases is not supported (READ_MAX_LEN)");
            if (d->meta.flags & SNK_TEXT_LEN_MISMATCH)
                die("sequence and quality have different lengths, read number " + std::to_string(d->first_index + d->meta.bad_record + 1));
            if (d->meta.flags & SNK_TEXT_LINE_COUNT) die("input fastq is truncated," + hp_.fq1_path);
            if (!(d->meta.flags & SNK_TEXT_STRIDE_OVERFLOW)) break;
            stride_ = std::max(stride_, round16(d->meta.max_len));       // a longer read showed up: wider rows, same batch again
            submit(d);
        }
        for (int m = 0; m < mates_; m++) {
            d->out[m].reserve((size_t)d->meta.out_bytes[m] + 64, 0);
            d->off[m].reserve(((size_t)d->n + 1) * sizeof(uint32_t), 0);
        }
        if (trim_) for (int m = 0; m < mates_; m++) d->res[m].reserve((size_t)d->n * sizeof(snk_read_result) + 64, 0);
        engine_check(snk_text_fetch_async(engines_[d->gpu], d->lane, d->out[0].p, pe_ ? d->out[1].p : nullptr, (uint32_t*)d->off[0].p,
                                          pe_ ? (uint32_t*)d->off[1].p : nullptr, trim_ ? (snk_read_result*)d->res[0].p : nullptr,
                                          (trim_ && pe_) ? (snk_read_result*)d->res[1].p : nullptr));
        engine_check(snk_engine_lane_sync(engines_[d->gpu], d->lane));
        t_gpu_wait_ += now_s() - t0;
        finish_batch(d);
    };
    HostBatch* b;
    while (gpu_q_.pop(b)) {
        if (inflight.size() >= depth) retire();
        submit(b);
        inflight.push_back(b);
    }
    while (!inflight.empty()) retire();
    tl_gpu_done_ = now_s() - t0_;
    { std::lock_guard<std::mutex> g(done_mu_); all_submitted_ = true; }
    done_cv_.notify_all();
    gz_q_.close();
}

void FilterRun::encode(const char* p, size_t n, std::string& out)
{
    // one gzip member per run. Default: the in-tree fast encoder (fast_deflate.h); SNK_GZ_CODEC=zlib: zlib level 2 as the
    // reference configures it (peprocess.cpp:1803-1810). Either way only the decompressed bytes are comparable.
    static const bool use_zlib = [] { const char* c = getenv("SNK_GZ_CODEC"); return c && strcmp(c, "zlib") == 0; }();
    if (!use_zlib) { out.clear(); fast_gzip_member((const uint8_t*)p, n, out); return; }
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, 2, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) die("zlib deflateInit2 failed");
    out.resize(deflateBound(&zs, (uLong)n) + 32);
    zs.next_in = (Bytef*)p; zs.avail_in = (uInt)n;
    zs.next_out = (Bytef*)&out[0]; zs.avail_out = (uInt)out.size();
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) die("zlib deflate failed");
    out.resize(zs.total_out);
    deflateEnd(&zs);
}

// Walks the batch's records in runs of one kind: 0 = in place, 1 = deferred, and markers (kind 2, r0 == r1) where
// the deferred bytes are emitted (see writer()). Within a cycle of cyc_ reads, [cyc_-defer_len_, cyc_) is
// deferred, and (from the second cycle on) the deferred bytes are emitted before read insert_off_.
template <class F>
void FilterRun::walk_runs(const HostBatch& b, F&& emit)
{
    if (!reorder_) { emit(0, 0u, b.n); return; }
    uint32_t r = 0;
    while (r < b.n) {
        const uint64_t gi = b.first_index + r, in_cyc = gi % cyc_;
        if (gi >= cyc_ && in_cyc == insert_off_) emit(2, r, r);
        const bool deferred = in_cyc >= cyc_ - defer_len_;
        uint64_t next = deferred ? cyc_ : cyc_ - defer_len_;          // next boundary inside the cycle
        if (gi >= cyc_ && in_cyc < insert_off_ && insert_off_ < next) next = insert_off_;
        else if (gi < cyc_ && !deferred) next = cyc_ - defer_len_;
        const uint64_t run = std::min<uint64_t>(next - in_cyc, b.n - r);
        emit(deferred ? 1 : 0, r, r + (uint32_t)run);
        r += (uint32_t)run;
    }
}

// Cuts the batch's clean text (device-formatted, input order) into pieces at the few record
// boundaries where the reference's emission order departs from input order (see writer()), and
// into <= 4 MiB runs for the parallel gzip members. The trim files (all records) get the same cuts
// as record ranges; their text is formatted by the workers.
void FilterRun::make_pieces(HostBatch& b)
{
    const size_t max_run = hp_.clean_gz ? (4u << 20) : ~(size_t)0;
    for (int m = 0; m < mates_; m++) {
        std::vector<Piece>& out = b.pieces[m];
        out.clear();
        const uint32_t* off = (const uint32_t*)b.off[m].p;
        const char* text = b.out[m].p;
        walk_runs(b, [&](int kind, uint32_t r0, uint32_t r1) {         // records [r0, r1)
            if (kind == 2) { out.push_back({2, nullptr, 0, std::string()}); return; }
            size_t a = off[r0];
            const size_t e = off[r1];
            while (a < e) {
                size_t stop = e;
                if (e - a > max_run) {                                     // cut at a record boundary near a + max_run
                    const uint32_t* it = std::upper_bound(off + r0, off + r1 + 1, (uint32_t)(a + max_run));
                    stop = (it == off + r0) ? e : (size_t)*(it - 1);
                    if (stop <= a) stop = (it == off + r1 + 1) ? e : (size_t)*it;
                }
                out.push_back({kind, text + a, stop - a, std::string()});
                if (kind == 1) {                                           // kept records of this run (warning text of writer())
                    const uint32_t* lo = std::lower_bound(off + r0, off + r1 + 1, (uint32_t)a);
                    const uint32_t* hi = std::lower_bound(off + r0, off + r1 + 1, (uint32_t)stop);
                    uint32_t k = 0;
                    for (const uint32_t* q = lo; q < hi; q++) k += q[1] != q[0];
                    out.back().kept = k;
                }
                a = stop;
            }
        });
        if (!trim_) continue;
        index_records(b, m);
        std::vector<Piece>& tout = b.tpieces[m];
        tout.clear();
        constexpr uint32_t kTrimRun = 8192;                                // records per gzip member
        walk_runs(b, [&](int kind, uint32_t r0, uint32_t r1) {
            if (kind == 2) { tout.push_back({2, nullptr, 0, std::string()}); return; }
            for (uint32_t a = r0; a < r1; a += kTrimRun) {
                Piece p{kind, nullptr, 0, std::string()};
                p.r0 = a; p.r1 = std::min(r1, a + kTrimRun);
                tout.push_back(std::move(p));
            }
        });
    }
}

// byte offset of every record of the mate's raw text (every 4th newline)
void FilterRun::index_records(HostBatch& b, int mate)
{
    std::vector<uint32_t>& rs = b.rec_start[mate];
    rs.resize((size_t)b.n + 1);
    const char* p = b.in[mate].p;
    const size_t end = b.in_bytes[mate];
    size_t pos = 0;
    for (uint32_t i = 0; i < b.n; i++) {
        rs[i] = (uint32_t)pos;
        size_t c = 0;
        pos += nth_newline(p + pos, end - pos, 4, &c);
    }
    rs[b.n] = (uint32_t)end;
}

// The trim files hold EVERY record as fastq_trim left it (peprocess.cpp:1460-1466, output_fastqs :3383-3433):
// id (index removal applied, one "/1" "/2" suffix with pe_info), the trimmed bases and qualities - possibly empty.
void FilterRun::format_trim(const HostBatch& b, int mate, uint32_t r0, uint32_t r1, std::string& out) const
{
    const char* text = b.in[mate].p;
    const snk_read_result* res = (const snk_read_result*)b.res[mate].p;
    const size_t strip = (size_t)fmt_.strip;
    const int qshift = hp_.out_quality_phred - hp_.quality_phred;
    out.clear();
    out.reserve((size_t)(r1 - r0) * (avg_rec_bytes_[mate] + 8));
    std::vector<uint8_t> idbuf;
    for (uint32_t r = r0; r < r1; r++) {
        const size_t rec_end = b.rec_start[mate][r + 1];
        size_t pos = b.rec_start[mate][r];
        const char* line[4]; size_t vis[4];
        for (int k = 0; k < 4; k++) {
            const char* nl = pos < rec_end ? (const char*)memchr(text + pos, '\n', rec_end - pos) : nullptr;
            const size_t raw = nl ? (size_t)(nl - (text + pos)) + 1 : rec_end - pos;
            line[k] = text + pos; vis[k] = raw > strip ? raw - strip : 0;
            pos += raw;
        }
        const size_t id0 = out.size();
        if (fmt_.id_mode == 0) out.append(line[0], vis[0]);
        else {
            idbuf.resize(vis[0] + 1);
            const uint32_t n = snkcore::id_transform((const uint8_t*)line[0], (uint32_t)vis[0], fmt_.id_mode, idbuf.data());
            out.append((const char*)idbuf.data(), n);
        }
        if (pe_ && hp_.pe_info) { out.push_back('/'); out.push_back(mate ? '2' : '1'); }
        if (fmt_.fasta) {
            const size_t at = out.find('@', id0);
            if (at != std::string::npos) out[at] = '>';
        }
        out.push_back('\n');
        const size_t h = res[r].head_cut, l = res[r].clean_len;
        out.append(line[1] + h, l);
        out.push_back('\n');
        if (!fmt_.fasta) {
            out.append("+\n", 2);
            const size_t q0 = out.size();
            out.append(line[3] + h, l);
            if (qshift) for (size_t k = q0; k < out.size(); k++) out[k] = (char)(out[k] + qshift);
            out.push_back('\n');
        }
    }
}

void FilterRun::finish_batch(HostBatch* b)
{
    make_pieces(*b);
    int tasks = 0;
    if (hp_.clean_gz)
        for (int m = 0; m < mates_; m++)
            for (const Piece& p : b->pieces[m]) tasks += p.kind != 2 && p.len > 0;
    if (trim_)
        for (int m = 0; m < mates_; m++)
            for (const Piece& p : b->tpieces[m]) tasks += p.kind != 2;
    if (tasks == 0) {
        { std::lock_guard<std::mutex> g(done_mu_); done_[b->seq_no] = b; }
        done_cv_.notify_all();
        return;
    }
    b->tasks = tasks;
    if (hp_.clean_gz)
        for (int m = 0; m < mates_; m++)
            for (size_t i = 0; i < b->pieces[m].size(); i++)
                if (b->pieces[m][i].kind != 2 && b->pieces[m][i].len > 0) gz_q_.push({b, m, i, false});
    if (trim_)
        for (int m = 0; m < mates_; m++)
            for (size_t i = 0; i < b->tpieces[m].size(); i++)
                if (b->tpieces[m][i].kind != 2) gz_q_.push({b, m, i, true});
}

void FilterRun::gz_worker()
{
    GzTask t;
    while (gz_q_.pop(t)) {
        NvtxRange nvtx("snk:gzip_member");
        const double t0 = now_s();
        Piece& p = t.trim ? t.b->tpieces[t.mate][t.piece] : t.b->pieces[t.mate][t.piece];
        if (t.trim) {
            std::string text;
            format_trim(*t.b, t.mate, p.r0, p.r1, text);
            encode(text.data(), text.size(), p.gz);
        } else encode(p.p, p.len, p.gz);
        p.p = p.gz.data(); p.len = p.gz.size();
        t_gz_us_ += (uint64_t)((now_s() - t0) * 1e6);
        if (--t.b->tasks == 0) {
            { std::lock_guard<std::mutex> g(done_mu_); done_[t.b->seq_no] = t.b; }
            done_cv_.notify_all();
        }
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
    // files 0,1 = clean fq1/fq2; files 2,3 = trim fq1/fq2 (trimFq1/2)
    int out[4] = {-1, -1, -1, -1};
    off_t pos[4] = {0, 0, 0, 0};
    const std::string names[4] = {hp_.output_dir + "/" + hp_.clean_fq1, hp_.output_dir + "/" + hp_.clean_fq2,
                                  hp_.output_dir + "/" + hp_.trim_fq1, hp_.output_dir + "/" + hp_.trim_fq2};
    const int nfiles = trim_ ? 4 : 2;
    bool via_mmap[4] = {false, false, false, false};
    off_t extended[4] = {0, 0, 0, 0};
    const long page = sysconf(_SC_PAGESIZE);
    for (int f = 0; f < nfiles; f++) {
        if (f % 2 >= mates_) continue;
        out[f] = open(names[f].c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
        if (out[f] < 0) out[f] = open(names[f].c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (out[f] < 0) die("cannot write to the file," + names[f]);
        // Several threads pwrite()-ing ONE file take turns on its inode lock, which caps a tmpfs output at a single
        // thread's page-cache copy rate. On tmpfs the pool therefore copies through shared mappings of the file (page
        // faults of different threads do not serialise); everywhere else it keeps pwrite (SNK_WRITE_MMAP=0/1 overrides).
        struct statfs sfs;
        struct stat st;
        const char* force = getenv("SNK_WRITE_MMAP");
        const bool regular = fstat(out[f], &st) == 0 && S_ISREG(st.st_mode) && (fcntl(out[f], F_GETFL) & O_ACCMODE) == O_RDWR;
        const bool tmpfs = fstatfs(out[f], &sfs) == 0 && (unsigned long)sfs.f_type == 0x01021994ul;
        via_mmap[f] = regular && (force ? atoi(force) != 0 : tmpfs);
    }
    // the order is fixed here (every run of bytes gets its file offset), the copying is done by a small pool
    struct WriteTask { int m; const char* p; size_t len; off_t at; HostBatch* owner; std::shared_ptr<std::string> hold; };
    Channel<WriteTask> tasks;
    auto write_at = [&](const WriteTask& t) {
        const char* p = t.p; size_t n = t.len; off_t at = t.at;
        if (via_mmap[t.m] && n >= (64u << 10)) {            // the file already extends to at + n (ftruncate by the ordering thread)
            const off_t a0 = at & ~(off_t)(page - 1);
            const size_t delta = (size_t)(at - a0);
            void* m = mmap(nullptr, n + delta, PROT_READ | PROT_WRITE, MAP_SHARED, out[t.m], a0);
            if (m != MAP_FAILED) {
                memcpy((char*)m + delta, p, n);
                munmap(m, n + delta);
                return;
            }
        }
        while (n > 0) {
            const ssize_t w = ::pwrite(out[t.m], p, std::min<size_t>(n, 1u << 30), at);
            if (w < 0) die("cannot write to the file," + names[t.m]);
            p += w; n -= (size_t)w; at += w;
        }
    };
    std::vector<std::thread> pool;
    int nwriters = std::max(2, std::min(8, hp_.threads / 2));
    if (const char* e = getenv("SNK_WRITE_THREADS")) nwriters = std::max(1, atoi(e));
    for (int i = 0; i < nwriters; i++)
        pool.emplace_back([&] {
            WriteTask t;
            while (tasks.pop(t)) {
                NvtxRange nvtx("snk:write");
                const double t0 = now_s();
                write_at(t);
                t_write_us_ += (uint64_t)((now_s() - t0) * 1e6);
                if (t.owner && --t.owner->tasks == 0) free_q_.push(t.owner);
            }
        });
    constexpr size_t kWriteRun = 8u << 20;
    uint64_t next = 0;
    std::vector<WriteTask> mine;
    for (;;) {
        HostBatch* b = nullptr;
        {
            std::unique_lock<std::mutex> g(done_mu_);
            done_cv_.wait(g, [&] { return done_.count(next) || (all_submitted_ && next >= n_batches_total_); });
            auto it = done_.find(next);
            if (it == done_.end()) break;
            b = it->second; done_.erase(it);
        }
        mine.clear();
        for (int f = 0; f < nfiles; f++) {
            if (f % 2 >= mates_) continue;
            for (Piece& p : (f < 2 ? b->pieces[f] : b->tpieces[f - 2])) {
                if (p.kind == 0) {
                    // large runs are written by several threads of the pool (page-cache copies again)
                    for (size_t a = 0; a < p.len; a += kWriteRun) mine.push_back({f, p.p + a, std::min(kWriteRun, p.len - a), pos[f] + (off_t)a, b, nullptr});
                    pos[f] += (off_t)p.len;
                }
                else if (p.kind == 1) { pending_deferred_[f].append(p.p, p.len); if (f == 0) deferred_records_ += p.kept; }
                else if (!pending_deferred_[f].empty()) {
                    if (f == 0) deferred_records_ = 0;
                    auto hold = std::make_shared<std::string>(std::move(pending_deferred_[f]));
                    pending_deferred_[f].clear();
                    mine.push_back({f, hold->data(), hold->size(), pos[f], nullptr, hold});
                    pos[f] += (off_t)hold->size();
                }
            }
        }
        if (b->seq_no % 16 == 0) log_line(local_time() + " processed_reads:\t" + std::to_string(b->first_index + b->n));
        next++;
        for (int f = 0; f < nfiles; f++)
            if (via_mmap[f] && pos[f] > extended[f]) {
                if (ftruncate(out[f], pos[f]) != 0) via_mmap[f] = false;
                else extended[f] = pos[f];
            }
        int owned = 0;
        for (const WriteTask& t : mine) owned += t.owner != nullptr;
        if (owned == 0) {
            for (const WriteTask& t : mine) tasks.push(t);
            for (int m = 0; m < mates_; m++) { b->pieces[m].clear(); b->tpieces[m].clear(); }
            free_q_.push(b);
        } else {
            b->tasks = owned;                 // the batch (its pinned text and gzip strings) is recycled by the last write
            for (const WriteTask& t : mine) tasks.push(t);
        }
    }
    tasks.close();
    for (auto& t : pool) t.join();
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
    if (drop && deferred_records_ > 0)
        std::cerr << "Warning:" << deferred_records_ << " clean read" << (pe_ ? " pairs" : "s") << " (the last deferred batch before the end of the input) are counted in the"
                  << " statistics but not written, exactly like SOAPnuke 2.1.9 loses them in its final concat pass (peprocess.cpp:2957-2966);"
                  << " set SNK_KEEP_DEFERRED=1 to write them" << std::endl;
    for (int f = 0; f < nfiles; f++) {
        if (f % 2 >= mates_) continue;
        if (!drop && !pending_deferred_[f].empty()) {
            if (via_mmap[f] && ftruncate(out[f], pos[f] + (off_t)pending_deferred_[f].size()) != 0) via_mmap[f] = false;
            write_at({f, pending_deferred_[f].data(), pending_deferred_[f].size(), pos[f], nullptr, nullptr});
        }
        if (close(out[f]) != 0) die("cannot write to the file," + names[f]);
    }
}

void FilterRun::process()
{
    const double t_begin = now_s();
    t0_ = t_begin;
    mkdir_p(hp_.output_dir);
    log_.open(hp_.log.c_str());
    if (!log_) die("cannot open such file," + hp_.log);
    log_line(local_time() + "\tAnalysis start!");
    if (snk_params_check(&ep_)) die(snk_last_error());
    if (hp_.output_file_type != "fasta" && hp_.output_file_type != "fastq") die("output_file_type value error");
    memset(&fmt_, 0, sizeof fmt_);
    fmt_.strip = 1;
    trim_ = !hp_.trim_fq1.empty();
    // seProcess::preOutput has no /1 (seprocess.cpp:919). With the trim files on, preOutput runs on the same record once
    // for the trim copy and once more for the clean copy (peprocess.cpp:1460-1475): the clean ids get the suffix twice.
    fmt_.pe_info = (pe_ && hp_.pe_info) ? (trim_ ? 2 : 1) : 0;
    fmt_.fasta = hp_.output_file_type == "fasta";
    fmt_.id_mode = hp_.index_remove ? (hp_.seq_type == "0" ? 1 : 2) : 0;
    // emission-order quirk applies to plain-text PE input with more than one worker
    cyc_ = (uint64_t)ep_.slot_block * (uint64_t)ep_.n_slots;
    defer_len_ = (uint64_t)hp_.patch_size;
    insert_off_ = (uint64_t)ep_.slot_block * (uint64_t)(ep_.n_slots - 1);
    reorder_ = pe_ && !hp_.input_gz && ep_.n_slots > 1;
    // The reader starts first: it opens the inputs and reads ahead while the CUDA contexts come up; it gets its first batch
    // buffer (pinned memory needs a context) only after that.
    std::thread t_ingest([&] { ingest(); });
    {
        // one engine per GPU; the CUDA contexts of different devices come up in parallel (0.3 - 1 s each)
        engines_.assign((size_t)hp_.n_gpus, nullptr);
        std::vector<std::string> errs((size_t)hp_.n_gpus);
        std::vector<std::thread> th;
        auto create = [&](int g) { if (snk_engine_create(&ep_, g, &engines_[g])) errs[g] = snk_last_error(); };   // the error text is per thread
        for (int g = 1; g < hp_.n_gpus; g++) th.emplace_back(create, g);
        create(0);
        for (auto& t : th) t.join();
        for (const std::string& e : errs) if (!e.empty()) die(e);
    }
    batches_.resize(inflight_depth() + 3);       // in flight on the GPUs + being read + being compressed/written
    for (auto& b : batches_) free_q_.push(&b);

    t_setup_ = now_s() - t_begin;
    const int nworkers = (hp_.clean_gz || trim_) ? std::max(2, hp_.threads) : 0;
    std::thread t_gpu([&] { gpu_stage(); });
    std::vector<std::thread> workers;
    for (int i = 0; i < nworkers; i++) workers.emplace_back([&] { gz_worker(); });
    std::thread t_writer([&] { writer(); });
    t_ingest.join();
    t_gpu.join();
    for (auto& w : workers) w.join();
    t_writer.join();
    tl_writer_done_ = now_s() - t0_;
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
        if (flags & 8) die("read longer than its batch row, read number " + std::to_string(bad + 1));
        engine_check(snk_engine_stats(e, part.data()));
        std::vector<uint64_t> keys;
        for (size_t k : key_words) { keys.push_back(std::max(total[k], part[k])); part[k] = 0; total[k] = 0; }
        for (size_t i = 0; i < words; i++) total[i] += part[i];
        for (size_t j = 0; j < key_words.size(); j++) total[key_words[j]] = keys[j];
    }
    tl_stats_done_ = now_s() - t0_;
    double stage_ms[SNK_STAGE_COUNT] = {0};
    for (snk_engine* e : engines_) {
        double ms[SNK_STAGE_COUNT];
        if (snk_engine_stage_times(e, ms) == 0) for (int i = 0; i < SNK_STAGE_COUNT; i++) stage_ms[i] += ms[i];
    }
    if (pe_) { if (snk_report_write_pe(&ep_, total.data(), hp_.output_dir.c_str())) die(snk_last_error()); }
    else { if (snk_report_write_se(&ep_, total.data(), hp_.output_dir.c_str())) die(snk_last_error()); }
    if (!hp_.fast_exit) {                        // the CLI leaves device and pinned memory to process exit
        for (snk_engine* e : engines_) snk_engine_destroy(e);
        engines_.clear();
        for (auto& b : batches_) b.release();
        batches_.clear();
    }
    {
        char buf[512];
        snprintf(buf, sizeof buf, "stage seconds: setup %.2f, read(busy) %.2f, gpu-wait %.2f, gzip(sum over %d workers) %.2f, write(sum over the writer pool) %.2f, total %.2f; reads %llu",
                 t_setup_, t_read_, t_gpu_wait_, nworkers, t_gz_us_.load() * 1e-6, t_write_us_.load() * 1e-6, now_s() - t_begin,
                 (unsigned long long)total_reads_);
        log_line(buf);
        snprintf(buf, sizeof buf, "device seconds (CUDA events, summed over %zu GPU(s) x lanes): h2d %.3f, line index + row packing %.3f, filter kernel %.3f, clean text formatting %.3f, d2h %.3f",
                 (size_t)hp_.n_gpus, stage_ms[SNK_STAGE_H2D] * 1e-3, stage_ms[SNK_STAGE_INDEX_PACK] * 1e-3, stage_ms[SNK_STAGE_FILTER] * 1e-3,
                 stage_ms[SNK_STAGE_FORMAT] * 1e-3, stage_ms[SNK_STAGE_D2H] * 1e-3);
        log_line(buf);
        snprintf(buf, sizeof buf, "timeline seconds since start: engines ready %.2f (%.2f GB of input read ahead meanwhile), first batch read %.2f, input exhausted %.2f, last batch off the GPU %.2f, outputs written %.2f, statistics gathered %.2f, reports written %.2f",
                 t_setup_, tl_prefetched_, tl_first_batch_, tl_ingest_done_, tl_gpu_done_, tl_writer_done_, tl_stats_done_, now_s() - t_begin);
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

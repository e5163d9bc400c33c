// gz_members.cpp — see gz_members.h
#include "gz_members.h"
#include <zlib.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace snk {

namespace {

constexpr size_t kChunk = 1u << 20;            // decoded bytes handed over at a time
constexpr size_t kMemberCap = 32u << 20;       // a member's worker pauses with this much decoded and unconsumed
constexpr size_t kBudget = 512u << 20;         // no new member is started (except the chain head) above this total
constexpr size_t kScanAhead = 1ull << 30;      // the scanner stays within this many compressed bytes of the chain head
constexpr int kClaimRun = 8;                   // consecutive small members a worker takes at once (BGZF blocks)
constexpr size_t kSmallMember = 256u << 10;    // "small": the next candidate is this close

struct Member {
    size_t start = 0, end = 0;
    enum State { Pending, Running, Done, Failed, Cancelled } st = Pending;
    std::deque<std::vector<char>> chunks;       // decoded, not yet consumed
    size_t buffered = 0;
};

inline bool header_at(const uint8_t* p, size_t left)
{
    return left >= 18 && p[0] == 0x1f && p[1] == 0x8b && p[2] == 8 && (p[3] & 0xE0) == 0;
}

} // namespace

struct GzMemberReader::Impl {
    int fd = -1;
    const uint8_t* data = nullptr;
    size_t size = 0;

    std::mutex mu;
    std::condition_variable cv_consumer, cv_workers;
    std::map<size_t, std::shared_ptr<Member>> members;   // by start offset
    size_t head = 0;             // chain position: start of the member the consumer is reading
    size_t head_off = 0;         // consumed bytes of the head's front chunk
    size_t claim_pos = 0;        // every candidate below this offset has been claimed
    size_t scan_pos = 0;         // candidates below this offset are all listed
    bool scan_done = false, stop = false, eof = false, error = false;
    size_t total_buffered = 0;
    Counters cnt;
    std::vector<std::thread> threads;

    void scanner();
    void worker();
    void inflate_member(const std::shared_ptr<Member>& m, z_stream& zs, std::vector<char>& scratch);
    void advance_head(std::unique_lock<std::mutex>& lk, size_t new_head);
};

// candidate member starts, in file order; paced by the chain head so that a huge file is not paged in far ahead
void GzMemberReader::Impl::scanner()
{
    size_t pos = 0;
    std::vector<size_t> found;
    while (pos < size) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_workers.wait(lk, [&] { return stop || pos < head + kScanAhead; });
            if (stop) return;
        }
        const size_t stop_at = std::min(size, pos + (4u << 20));
        found.clear();
        size_t i = pos;
        while (i < stop_at) {
            const uint8_t* q = (const uint8_t*)memchr(data + i, 0x1f, stop_at - i);
            if (!q) break;
            i = (size_t)(q - data);
            if (header_at(q, size - i)) found.push_back(i);
            i++;
        }
        pos = stop_at;
        {
            std::lock_guard<std::mutex> g(mu);
            for (size_t s : found)
                if (s >= head && !members.count(s)) { auto m = std::make_shared<Member>(); m->start = s; members[s] = m; }
            scan_pos = pos;
        }
        cv_workers.notify_all();
        cv_consumer.notify_all();
    }
    { std::lock_guard<std::mutex> g(mu); scan_done = true; scan_pos = size; }
    cv_workers.notify_all();
    cv_consumer.notify_all();
}

void GzMemberReader::Impl::inflate_member(const std::shared_ptr<Member>& m, z_stream& zs, std::vector<char>& scratch)
{
    inflateReset(&zs);
    const uint8_t* in = data + m->start;
    size_t in_left = size - m->start;
    zs.next_in = (Bytef*)in; zs.avail_in = 0;
    auto finish = [&](Member::State st, size_t end) {
        std::lock_guard<std::mutex> g(mu);
        if (m->st == Member::Running) { m->st = st; m->end = end; }
        if (m->start == head) cv_consumer.notify_all();
    };
    for (;;) {
        if (zs.avail_in == 0 && in_left > 0) {
            const size_t feed = std::min<size_t>(in_left, 1u << 30);
            zs.avail_in = (uInt)feed; in_left -= feed;
        }
        zs.next_out = (Bytef*)scratch.data(); zs.avail_out = (uInt)scratch.size();
        const int rc = inflate(&zs, Z_NO_FLUSH);
        const size_t produced = scratch.size() - zs.avail_out;
        if (produced) {
            std::vector<char> chunk(scratch.data(), scratch.data() + produced);
            std::unique_lock<std::mutex> lk(mu);
            if (m->st != Member::Running) return;                       // cancelled: not on the chain
            m->chunks.push_back(std::move(chunk));
            m->buffered += produced; total_buffered += produced;
            if (m->start == head) cv_consumer.notify_all();
            cv_workers.wait(lk, [&] { return stop || m->st != Member::Running || m->buffered < kMemberCap; });
            if (stop || m->st != Member::Running) return;
        }
        if (rc == Z_STREAM_END) { finish(Member::Done, (size_t)((const uint8_t*)zs.next_in - data)); return; }
        if (rc == Z_OK) continue;
        if (rc == Z_BUF_ERROR && (zs.avail_in > 0 || in_left > 0 || zs.avail_out == 0)) continue;   // wants more room or input that exists
        finish(Member::Failed, 0);                                      // corrupt data, or the file ends inside the member
        return;
    }
}

void GzMemberReader::Impl::worker()
{
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, 15 + 16) != Z_OK) { std::lock_guard<std::mutex> g(mu); error = true; cv_consumer.notify_all(); return; }
    std::vector<char> scratch(kChunk);
    std::vector<std::shared_ptr<Member>> run;
    for (;;) {
        run.clear();
        {
            std::unique_lock<std::mutex> lk(mu);
            for (;;) {
                if (stop) { inflateEnd(&zs); return; }
                if (claim_pos < head) claim_pos = head;
                auto it = members.lower_bound(claim_pos);
                while (it != members.end() && it->second->st != Member::Pending) ++it;
                if (it != members.end() && (it->first == head || total_buffered < kBudget)) {
                    // a run of consecutive small candidates, so that 64 KiB BGZF blocks do not pay a lock round trip each
                    for (int k = 0; k < kClaimRun && it != members.end() && it->second->st == Member::Pending; k++) {
                        it->second->st = Member::Running;
                        run.push_back(it->second);
                        claim_pos = it->first + 1;
                        auto nx = std::next(it);
                        if (nx == members.end() || nx->first - it->first > kSmallMember) break;
                        it = nx;
                    }
                    break;
                }
                cv_workers.wait(lk);
            }
        }
        for (auto& m : run) inflate_member(m, zs, scratch);
    }
}

// the chain moved on: everything that starts before the new head is either consumed or was never a member
void GzMemberReader::Impl::advance_head(std::unique_lock<std::mutex>&, size_t new_head)
{
    for (auto it = members.begin(); it != members.end() && it->first < new_head;) {
        Member& m = *it->second;
        if (it->first != head) { cnt.cancelled++; cnt.bytes_dropped += m.buffered; }
        total_buffered -= m.buffered;
        m.buffered = 0; m.chunks.clear();
        m.st = Member::Cancelled;
        it = members.erase(it);
    }
    head = new_head; head_off = 0;
    cv_workers.notify_all();
}

GzMemberReader* GzMemberReader::open(const std::string& path, int nthreads)
{
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return nullptr;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 18) { ::close(fd); return nullptr; }
    void* p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (p == MAP_FAILED) { ::close(fd); return nullptr; }
    if (!header_at((const uint8_t*)p, (size_t)st.st_size)) { munmap(p, (size_t)st.st_size); ::close(fd); return nullptr; }
    madvise(p, (size_t)st.st_size, MADV_SEQUENTIAL);
    GzMemberReader* r = new GzMemberReader();
    Impl* d = r->d_ = new Impl();
    d->fd = fd; d->data = (const uint8_t*)p; d->size = (size_t)st.st_size;
    { auto m = std::make_shared<Member>(); m->start = 0; d->members[0] = m; }
    d->threads.emplace_back([d] { d->scanner(); });
    for (int i = 0; i < std::max(1, nthreads); i++) d->threads.emplace_back([d] { d->worker(); });
    return r;
}

GzMemberReader::~GzMemberReader()
{
    if (!d_) return;
    { std::lock_guard<std::mutex> g(d_->mu); d_->stop = true; }
    d_->cv_workers.notify_all();
    d_->cv_consumer.notify_all();
    for (auto& t : d_->threads) t.join();
    munmap((void*)d_->data, d_->size);
    ::close(d_->fd);
    delete d_;
}

GzMemberReader::Counters GzMemberReader::counters() const
{
    std::lock_guard<std::mutex> g(d_->mu);
    return d_->cnt;
}

size_t GzMemberReader::read(char* dst, size_t n)
{
    Impl& d = *d_;
    size_t got = 0;
    std::unique_lock<std::mutex> lk(d.mu);
    while (got < n) {
        if (d.error) return kError;
        if (d.eof) break;
        auto it = d.members.find(d.head);
        if (it == d.members.end()) {
            // gzread(): anything that is not a gzip header behind a complete member is trailing garbage and ignored
            if (d.head + 2 > d.size || d.data[d.head] != 0x1f || d.data[d.head + 1] != 0x8b) { d.eof = true; break; }
            if (!d.scan_done && d.scan_pos <= d.head) { if (got) break; d.cv_consumer.wait(lk); continue; }
            // a gzip magic the scanner did not list (unusual flag bits): let zlib judge the header
            auto m = std::make_shared<Member>(); m->start = d.head; d.members[d.head] = m;
            if (d.claim_pos > d.head) d.claim_pos = d.head;
            d.cv_workers.notify_all();
            continue;
        }
        Member& m = *it->second;
        if (!m.chunks.empty()) {
            std::vector<char>& c = m.chunks.front();
            const size_t take = std::min(n - got, c.size() - d.head_off);
            const char* src = c.data() + d.head_off;
            const bool whole = d.head_off + take == c.size();
            if (take >= (64u << 10)) {                    // copy outside the lock; the chunk is only touched by the consumer
                lk.unlock();
                memcpy(dst + got, src, take);
                lk.lock();
            } else memcpy(dst + got, src, take);
            got += take; d.head_off += take;
            // wake the workers only when this frees something they can be waiting for
            const bool wake = m.buffered >= kMemberCap || d.total_buffered >= kBudget;
            m.buffered -= take; d.total_buffered -= take; d.cnt.bytes_out += take;
            if (whole) { m.chunks.pop_front(); d.head_off = 0; }
            if (wake) d.cv_workers.notify_all();
            continue;
        }
        if (m.st == Member::Done) { d.cnt.members++; d.advance_head(lk, m.end); continue; }
        if (m.st == Member::Failed) {
            // the very first member decides "is this gzip at all"; later ones: gzread() reports the error
            d.error = true;
            return got ? got : kError;
        }
        if (got) break;                                   // hand over what is there instead of waiting for more
        d.cv_consumer.wait(lk);
    }
    return got;
}

} // namespace snk

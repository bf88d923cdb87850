// Sample sink: queue-backed CSV writer with the observable behaviour of the reference's
// moodycamel::ConcurrentQueue<Eigen::VectorXd> + ofstream pair (reference src/BayesRv2.cpp:62,72,260-267,282-289;
// same pattern in the other three samplers): the sampler thread enqueue()s one row per kept iteration, a consumer
// thread try_dequeue()s and writes `row.transpose().format(CommaInitFmt)` -- "%g"-formatted numbers (ostream
// precision 6) joined by ", " -- plus '\n'.  One producer, one consumer, so a bounded ring replaces the MPMC queue.
// Deliberate differences: the stop flag is atomic and the consumer DRAINS the queue after stop,
// so every kept row reaches the file, in order, for all four samplers.
#include "common.cuh"
#include "writer.h"
#include <algorithm>
#include <cstdio>
#include <thread>
#include <vector>
#include <cstring>
#include <charconv>
#include <cmath>

namespace brr {

RowQueue::RowQueue(size_t capacity) : slots_(capacity + 1) {}

bool RowQueue::try_enqueue(std::vector<double> &&row)
{
    const size_t h = head_.load(std::memory_order_relaxed), n = (h + 1) % slots_.size();
    if (n == tail_.load(std::memory_order_acquire)) return false;   // full
    slots_[h] = std::move(row);
    head_.store(n, std::memory_order_release);
    return true;
}
void RowQueue::enqueue(std::vector<double> &&row)
{
    std::vector<double> r = std::move(row);
    while (true) {
        const size_t h = head_.load(std::memory_order_relaxed), n = (h + 1) % slots_.size();
        if (n != tail_.load(std::memory_order_acquire)) { slots_[h] = std::move(r); head_.store(n, std::memory_order_release); return; }
        std::this_thread::yield();
    }
}
bool RowQueue::try_dequeue(std::vector<double> &row)
{
    const size_t t = tail_.load(std::memory_order_relaxed);
    if (t == head_.load(std::memory_order_acquire)) return false;   // empty
    row = std::move(slots_[t]);
    tail_.store((t + 1) % slots_.size(), std::memory_order_release);
    return true;
}

static void put_indexed(std::string &s, const char *name, long i, const char *tail)
{
    char t[24];
    const auto r = std::to_chars(t, t + sizeof t, i);
    s += name; s.push_back('['); s.append(t, (size_t)(r.ptr - t)); s.push_back(']'); s += tail;
}

// Header text per sampler: src/BayesRv2.cpp:16-37, src/BayesRv2Groups.cpp:25-54, src/HorseshoeR.cpp:279-291 (trailing comma).
// BRV2Grstart defines a header writer (src/BRv2Grstart.cpp:26-50) but never calls it: its file has no header line.
std::string sample_header(int kind, int64_t N, int64_t M, int G, int64_t F)
{
    std::string s;
    if (kind == BRR_GRSTART) return s;
    s.reserve((size_t)(2 * M + N) * 14);
    s += "iteration,"; s += "mu,";
    for (int64_t i = 0; i < M; ++i) put_indexed(s, "beta", i + 1, ",");
    if (kind == BRR_V2) {
        s += "sigmaE,"; s += "sigmaG,";
        for (int64_t i = 0; i < M; ++i) put_indexed(s, "comp", i + 1, ",");
        for (int64_t i = 0; i < N - 1; ++i) put_indexed(s, "epsilon", i + 1, ",");
        put_indexed(s, "epsilon", N, "");
    } else if (kind == BRR_GROUPS) {
        s += "sigmaE,";
        for (int64_t i = 0; i < M; ++i) put_indexed(s, "comp", i + 1, ",");
        for (int i = 0; i < G; ++i) put_indexed(s, "sigmaG", i + 1, ",");
        for (int64_t i = 0; i < N - 1; ++i) put_indexed(s, "epsilon", i + 1, ",");
        put_indexed(s, "epsilon", N, ",");
        for (int64_t i = 0; i < F; ++i) put_indexed(s, "alpha", i + 1, ",");
        s += "sigmaF";
    } else {
        s += "sigmaE,"; s += "tau,";
        for (int64_t i = 0; i < M; ++i) put_indexed(s, "lambda", i + 1, ",");
        for (int64_t i = 0; i < N; ++i) put_indexed(s, "epsilon", i + 1, ",");
    }
    s += "\n";
    return s;
}

// "%g" of every value (ostream precision 6, reference src/BayesRv2.cpp:72), joined by ", ".  Rows are long (2M + 4 + N numbers)
// and mostly zeros and small integers, so those take a short cut; everything else goes through std::to_chars, which is
// specified to produce what printf("%.6g") produces in the C locale, without the locale / format-string machinery.
// values [0, len) of a span; `lead`: the span continues a row (separator before its first value too)
static void format_span(const double *row, size_t len, bool lead, std::string &out)
{
    out.clear();
    out.reserve(len * 12 + 2);
    char t[48];
    for (size_t i = 0; i < len; ++i) {
        const double v = row[i];
        if (i || lead) out.append(", ", 2);
        if (v == 0.0 && !std::signbit(v)) { out.push_back('0'); continue; }
        if (v > 0.0 && v < 100000.0 && v == (double)(int)v) {          // "%g" prints integers below 10^6 as plain digits
            int k = (int)v, n = 0;
            char d[8];
            while (k) { d[n++] = (char)('0' + k % 10); k /= 10; }
            while (n) out.push_back(d[--n]);
            continue;
        }
        if (std::isfinite(v)) {
            const auto r = std::to_chars(t, t + sizeof t, v, std::chars_format::general, 6);
            out.append(t, (size_t)(r.ptr - t));
        } else {
            const int n = snprintf(t, sizeof t, "%g", v);                 // nan / inf spellings as printf has them
            out.append(t, (size_t)n);
        }
    }
}
void format_row(const double *row, size_t len, std::string &out)
{
    format_span(row, len, false, out);
    out.push_back('\n');
}

std::string binary_sample_header(int kind, int64_t N, int64_t M, int G, int64_t F, int64_t row_len)
{
    std::string h(64, '\0');
    memcpy(&h[0], "BRRSMP1", 7);
    const int32_t k = kind, g = G;
    memcpy(&h[8], &k, 4); memcpy(&h[12], &g, 4); memcpy(&h[16], &N, 8); memcpy(&h[24], &M, 8); memcpy(&h[32], &F, 8); memcpy(&h[40], &row_len, 8);
    return h;
}

void SampleWriter::write_row(const std::vector<double> &row, std::string &text)
{
    if (binary_) {
        if (fwrite(row.data(), 8, row.size(), f_) != row.size()) io_error_.store(true);
    } else if (row.size() < (size_t)1 << 16) {
        format_row(row.data(), row.size(), text);
        if (fwrite(text.data(), 1, text.size(), f_) != text.size()) io_error_.store(true);
    } else {
        // long rows (a row carries the residual of every individual: 2M + 4 + N numbers) are formatted by several threads, span by
        // span, and written in order: at 400,000 individuals one thread needs ~7 ms per row and the chain produces one every ~20 ms
        const size_t hw = std::max(2u, std::thread::hardware_concurrency());
        const size_t T = std::min<size_t>({ (size_t)8, hw - 1, row.size() >> 15 });
        const size_t per = (row.size() + T - 1) / T;
        std::vector<std::string> parts(T);
        std::vector<std::thread> th;
        for (size_t t = 1; t < T; ++t)
            th.emplace_back([&, t] { const size_t lo = t * per, hi = std::min(row.size(), lo + per); if (lo < hi) format_span(row.data() + lo, hi - lo, true, parts[t]); });
        format_span(row.data(), std::min(per, row.size()), false, parts[0]);
        for (auto &x : th) x.join();
        parts[T - 1].push_back('\n');
        for (size_t t = 0; t < T; ++t)
            if (fwrite(parts[t].data(), 1, parts[t].size(), f_) != parts[t].size()) io_error_.store(true);
    }
    rows_written_.fetch_add(1);
}

SampleWriter::SampleWriter(const std::string &path, const std::string &header, bool write_header_now, bool binary)
    : q_(8), header_(header), binary_(binary)
{
    f_ = fopen(path.c_str(), binary ? "wb" : "w");   // truncates like ofstream::open (src/BayesRv2.cpp:69)
    if (!f_) throw Error(BRR_E_IO, "cannot open output file '" + path + "'");
    if (write_header_now && !header_.empty()) { fwrite(header_.data(), 1, header_.size(), f_); header_.clear(); }
}
void SampleWriter::start()
{
    if (running_) return;
    running_ = true;
    stop_.store(false);
    th_ = std::thread([this] {
        if (!header_.empty()) { fwrite(header_.data(), 1, header_.size(), f_); header_.clear(); }
        std::vector<double> row; std::string text;
        while (true) {
            if (q_.try_dequeue(row)) {
                write_row(row, text);
            } else if (stop_.load(std::memory_order_acquire)) {
                if (!q_.try_dequeue(row)) break;          // drained
                write_row(row, text);
            } else {
                std::this_thread::yield();
            }
        }
    });
}
void SampleWriter::enqueue(const double *row, size_t len)
{
    if (!running_) start();
    q_.enqueue(std::vector<double>(row, row + len));
}
void SampleWriter::finish()
{
    if (running_) {
        stop_.store(true, std::memory_order_release);
        th_.join();
        running_ = false;
    }
    if (f_) {
        if (!header_.empty()) { fwrite(header_.data(), 1, header_.size(), f_); header_.clear(); }
        if (fflush(f_) != 0) io_error_.store(true);
        fclose(f_); f_ = nullptr;
    }
    if (io_error_.load()) throw Error(BRR_E_IO, "write to the sample file failed");
}
SampleWriter::~SampleWriter()
{
    try { finish(); } catch (...) { }
}

}  // namespace brr

// exposed for hosts and tests: the text of one sample row exactly as the writer emits it (without the trailing newline)
extern "C" int64_t brr_format_row(const double *row, int64_t len, char *out, int64_t cap)
{
    if (!row || len < 0) return -1;
    std::string text;
    brr::format_row(row, (size_t)len, text);
    const int64_t n = (int64_t)text.size() - 1;
    if (out && cap > 0) { const int64_t c = n < cap - 1 ? n : cap - 1; memcpy(out, text.data(), (size_t)c); out[c] = 0; }
    return n;
}

// test hook (CPU only): `nrows` rows through the queue-backed writer of the product -- value 0 of every row replaced by its index --
// into `path`, CSV text or raw fp64.  Used by tests/test_abi.py (long rows take the multi-threaded formatting path) and by the
// ThreadSanitizer build of this file (tools/tsan_writer.sh).
extern "C" int brr_writer_selftest(const char *path, const double *row, int64_t len, int64_t nrows, int binary)
{
    if (!path || !row || len < 1 || nrows < 0) return BRR_E_ARG;
    try {
        brr::SampleWriter w(path, std::string(), true, binary != 0);
        std::vector<double> r(row, row + len);
        for (int64_t i = 0; i < nrows; ++i) { r[0] = (double)i; w.enqueue(r.data(), r.size()); }
        w.finish();
        return w.rows_written() == (uint64_t)nrows ? BRR_OK : BRR_E_IO;
    } catch (const std::exception &) { return BRR_E_IO; }
}

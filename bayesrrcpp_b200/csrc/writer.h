// Queue-backed sample writer (see writer.cpp).
#pragma once
#include <atomic>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>
#include <stdexcept>
#include "../../include/bayesrr_b200.h"

namespace brr {

struct Error;   // common.cuh

// Bounded single-producer / single-consumer ring with the enqueue / try_dequeue surface the reference uses
// (reference src/concurrentqueue.h:911,1041 as called from src/BayesRv2.cpp:261,286).
class RowQueue {
public:
    explicit RowQueue(size_t capacity);
    bool try_enqueue(std::vector<double> &&row);
    void enqueue(std::vector<double> &&row);          // waits (yielding) while the ring is full
    bool try_dequeue(std::vector<double> &row);
private:
    std::vector<std::vector<double>> slots_;
    std::atomic<size_t> head_{0}, tail_{0};
};

std::string sample_header(int kind, int64_t N, int64_t M, int G, int64_t F);
void format_row(const double *row, size_t len, std::string &out);

class SampleWriter {
public:
    // write_header_now: V2 / Groups write the header before sampling starts (reference src/BayesRv2.cpp:70);
    // Horseshoe writes it from the consumer (src/HorseshoeR.cpp:275-291) -- same bytes either way.
    // binary: rows are written as raw little-endian fp64 (lossless; `header` is then the binary file header, brr_sample_file_header)
    SampleWriter(const std::string &path, const std::string &header, bool write_header_now, bool binary = false);
    ~SampleWriter();
    void start();
    void enqueue(const double *row, size_t len);
    void finish();                                    // stop signal, drain, join, close; throws on I/O error
    uint64_t rows_written() const { return rows_written_.load(); }
private:
    RowQueue q_;
    std::string header_;
    FILE *f_ = nullptr;
    std::thread th_;
    std::atomic<bool> stop_{false}, io_error_{false};
    std::atomic<uint64_t> rows_written_{0};
    bool running_ = false, binary_ = false;
    void write_row(const std::vector<double> &row, std::string &text);
};

// 64-byte header of the binary sample file: "BRRSMP1\0", int32 kind, int32 groups, int64 N, int64 M, int64 F, int64 row_len, zero padding
std::string binary_sample_header(int kind, int64_t N, int64_t M, int G, int64_t F, int64_t row_len);

}  // namespace brr

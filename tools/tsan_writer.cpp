// ThreadSanitizer harness of the queue-backed sample writer (bayesrrcpp_b200/csrc/writer.cpp, compiled as plain C++): an enqueue
// storm of short rows through the bounded ring (capacity 8) and a few long rows through the multi-threaded formatting path.
// Build + run: tools/tsan_writer.sh
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
extern "C" int brr_writer_selftest(const char *path, const double *row, int64_t len, int64_t nrows, int binary);
namespace brr { void set_last_error(const std::string &) {} }   // lives in geno.cu in the product
int main()
{
    std::vector<double> small(64), big(150000);
    for (size_t i = 0; i < small.size(); ++i) small[i] = (double)i * 0.25 - 3.0;
    for (size_t i = 0; i < big.size(); ++i) big[i] = (i % 3) ? 0.0 : (double)i * 1e-3;
    int rc = brr_writer_selftest("/tmp/brr_tsan_small.csv", small.data(), (int64_t)small.size(), 1000, 0);
    rc |= brr_writer_selftest("/tmp/brr_tsan_small.bin", small.data(), (int64_t)small.size(), 1000, 1);
    rc |= brr_writer_selftest("/tmp/brr_tsan_big.csv", big.data(), (int64_t)big.size(), 6, 0);
    printf("writer self-test under ThreadSanitizer: rc=%d\n", rc);
    return rc;
}

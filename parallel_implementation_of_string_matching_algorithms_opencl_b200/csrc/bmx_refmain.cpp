// bmx_refmain.cpp -- the reference's console program on top of libbmx.so.
//
// Mirrors the flow of BoyreMoore/BoyreMoore/BoyreMoore.cpp main() (:70-316) so that a user of the
// reference finds the same inputs, the same stages and the same console lines:
//   reads  inputEd.txt / input1Search.txt from the working directory           (:77-84)
//   splits the text into numberOfProcesses = 2 word ranges                      (:72, :94-141)
//   builds badSymTab / goodSymTab                                               (:150-190)
//   runs the search 10 times and averages the time                              (:211, :314-315)
//   prints "Found by <id> at : <pos>" per match and                            (kernel1.cl:24)
//          "The no. of occurrences by process <id> is <count>" per range        (:294-295)
// The OpenCL layer (:192-313) is replaced by bmx_search_partitions / bmx_search; the kernel file
// kernel1.cl is not needed.  After the reference-compatible part it also prints the SERIAL result
// (one range, what the north star pins parity on), which additionally finds the matches the word
// partition loses at the seam.
//
//   bmx_refmain [text-file [pattern-file [numberOfProcesses]]] [--quiet]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "bmx.h"

static std::string slurp(const char *path)
{
    std::ifstream ifs(path, std::ios::binary);
    return std::string((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
}

int main(int argc, char **argv)
{
    const char *text_path = "inputEd.txt", *pat_path = "input1Search.txt";
    int numberOfProcesses = 2;
    bool quiet = false;
    int npos = 0;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--quiet")) quiet = true;
        else if (npos == 0) { text_path = argv[i]; ++npos; }
        else if (npos == 1) { pat_path = argv[i]; ++npos; }
        else if (npos == 2) { numberOfProcesses = atoi(argv[i]); ++npos; }
    }
    const std::string text = slurp(text_path), pattern = slurp(pat_path);
    const int64_t n = (int64_t)text.size();
    const int32_t m = (int32_t)pattern.size();
    if (m <= 0 || numberOfProcesses <= 0) {
        fprintf(stderr, "empty pattern (%s) or bad process count\n", pat_path);
        return 1;
    }
    if (!quiet) fwrite(text.data(), 1, text.size(), stdout);   // the reference echoes the text (:92)

    std::vector<int32_t> start_endi(2 * (size_t)numberOfProcesses), ans((size_t)numberOfProcesses);
    std::vector<int32_t> goodSymTab((size_t)m);
    int32_t badSymTab[256];
    if (bmx_partition_words(text.data(), n, numberOfProcesses, start_endi.data()) != BMX_OK ||
        bmx_build_tables(pattern.data(), m, badSymTab, goodSymTab.data()) != BMX_OK) {
        fprintf(stderr, "bmx: %s\n", bmx_last_error());
        return 1;
    }

    double total_time = 0;
    for (int testNum = 0; testNum < 10; ++testNum) {
        const auto begin = std::chrono::steady_clock::now();
        if (bmx_search_partitions(text.data(), pattern.data(), start_endi.data(), ans.data(), goodSymTab.data(),
                                  badSymTab, m, numberOfProcesses) != BMX_OK) {
            fprintf(stderr, "bmx: %s\n", bmx_last_error());
            return 1;
        }
        const double time_spent = std::chrono::duration<double>(std::chrono::steady_clock::now() - begin).count();
        total_time += time_spent;
        if (testNum == 0 && !quiet) {
            // the kernel's device printf, reproduced from the position lists of each range
            for (int id = 0; id < numberOfProcesses; ++id) {
                const int64_t lo = start_endi[2 * id], len = (int64_t)start_endi[2 * id + 1] - lo + 1;
                if (len < m) continue;
                std::vector<int64_t> pos((size_t)(len - m + 1));
                uint64_t cnt = 0;
                if (bmx_search(text.data() + lo, len, pattern.data(), m, pos.data(), (int64_t)pos.size(), &cnt) != BMX_OK) {
                    fprintf(stderr, "bmx: %s\n", bmx_last_error());
                    return 1;
                }
                for (uint64_t i = 0; i < cnt; ++i) printf("\nFound by %d at : %lld", id, (long long)(pos[i] + lo));
            }
        }
        for (int x = 0; x < numberOfProcesses; ++x) printf("\nThe no. of occurrences by process %d is %d", x, ans[x]);
        printf("Time Spent:%g", time_spent);
    }
    printf("Average time  = %g\n", total_time / 10);

    uint64_t serial = 0;
    if (bmx_search(text.data(), n, pattern.data(), m, nullptr, 0, &serial) != BMX_OK) {
        fprintf(stderr, "bmx: %s\n", bmx_last_error());
        return 1;
    }
    long long partitioned = 0;
    for (int x = 0; x < numberOfProcesses; ++x) partitioned += ans[x];
    printf("Serial result (one range, (m-1)-byte halos): %llu occurrences; the %d-way word partition reports %lld\n",
           (unsigned long long)serial, numberOfProcesses, partitioned);
    return 0;
}

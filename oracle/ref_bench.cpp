// TEST INFRASTRUCTURE (CPU baseline driver) - not part of the product.
//
// Runs the UNMODIFIED reference (oracle/_ref/libspiral_ref_<cfg>_<isa>.so, built from
// /root/reference by oracle/Makefile) through its own harness and stock code path:
//   1. the reference's own `main` (src/spiral.cpp:1228) does argument parsing, table setup,
//      load_db() and one full do_test() query (client + server + "Is correct?" check);
//   2. every further repetition calls the reference's `do_test()` (src/spiral.cpp:2408) again
//      on the same database, so the ~10 s DB generation is paid once.
// The reference prints its own per-stage timers (print_summary, src/spiral.cpp:209-265) after
// every query; bench.py parses those lines.  Nothing here touches the reference's arithmetic.
//
// usage: ref_bench <nu_1> <nu_2> <idx> <queries> [extra reference flags...]
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>

void do_test();   // defined by the reference library (src/spiral.cpp:2408)

int main(int argc, char **argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: %s nu_1 nu_2 idx queries [reference flags]\n", argv[0]);
        return 2;
    }
    typedef int (*main_fn)(int, char **);
    main_fn ref_main = (main_fn)dlsym(RTLD_NEXT, "main");
    if (!ref_main) { fprintf(stderr, "reference main not found: %s\n", dlerror()); return 3; }
    int queries = atoi(argv[4]);

    std::vector<char *> a;
    a.push_back(argv[0]); a.push_back(argv[1]); a.push_back(argv[2]); a.push_back(argv[3]);
    if (argc > 5) {                       // the reference only parses flags from argv[5] on
        static char dummy[] = "a";
        a.push_back(dummy);
        for (int i = 5; i < argc; i++) a.push_back(argv[i]);
    }
    printf("=== ref_bench query 0 ===\n"); fflush(stdout);
    ref_main((int)a.size(), a.data());    // falls off the end without `return` - legal only for main
    for (int q = 1; q < queries; q++) {
        printf("=== ref_bench query %d ===\n", q); fflush(stdout);
        do_test();
    }
    fflush(stdout);
    return 0;
}

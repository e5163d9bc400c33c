// main.cpp — `SOAPnuke <module> [options]` (reference main.cpp:17-68): module dispatch, parameter
// parsing, then peProcess / seProcess ::process(). The `filter` module (and its alias `filterMeta`,
// process_argv.cpp:27,763) and `filtersRNA` (seProcess with the sRNA branches) are served by the GPU engine.
#include <iostream>
#include <string>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include "cli_params.h"
#include "process.h"

// SNK_TIMESTAMPS=1: wall-clock marks on stderr, to separate process start / driver teardown from the run itself
static void timestamp(const char* what)
{
    if (!getenv("SNK_TIMESTAMPS")) return;
    struct timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    fprintf(stderr, "snk-ts %s %.6f\n", what, ts.tv_sec + 1e-9 * ts.tv_nsec);
}

int main(int argc, char** argv)
{
    timestamp("main");
    if (argc < 2) {
        std::cout << "\nProgram: SOAPnuke (b200 filter engine)\nCommand:\n         filter        preprocessing normal Fastq files\n"
                     "         filterMeta    preprocessing Meta Fastq files\n         filtersRNA    preprocessing sRNA Fastq files\n\n";
        return 1;
    }
    const std::string module = argv[1];
    if (module == "-h" || module == "--help") { snk::print_usage("filter"); return 0; }
    if (module == "-v" || module == "--version") { snk::print_version(); return 0; }
    if (module != "filter" && module != "filterMeta" && module != "filtersRNA") {
        if (module == "filterStLFR" || module == "filterHts")
            std::cerr << "Error:module " << module << " is not served by the GPU filter engine (only filter / filterMeta / filtersRNA)" << std::endl;
        else
            std::cerr << "Error:no such module,type -h/--help for help" << std::endl;      // process_argv.cpp:16-36
        return 1;
    }
    snk::HostParams hp;
    if (snk::parse_command_line(argc, argv, hp)) return 0;
    hp.fast_exit = true;
    if (hp.is_pe) { snk::peProcess* p = new snk::peProcess(hp); p->process(); }
    else { snk::seProcess* p = new snk::seProcess(hp); p->process(); }
    // every output file is written and closed: skip the CUDA context / pinned memory teardown
    std::cout.flush();
    std::cerr.flush();
    timestamp("exit");
    _exit(0);
}

// process.h — peProcess / seProcess: the entry points main() calls (reference main.cpp:57-65,
// peprocess.h:56-57, seprocess.h:34-35), re-designed as a pinned-buffer batching driver:
//
//   reader threads (gz/plain decode + FASTQ split + fixed-stride SoA pack into pinned memory)
//     -> engine lanes (cudaMemcpyAsync H2D, filter_kernel, D2H of the 8-byte result records)
//     -> worker pool (format surviving records, gzip level-2 members)
//     -> ordered writer (reference output order, see Writer::route)
//   then statistics gather across GPUs + report writer.
//
// The four reference calls this replaces per batch - filter_pe_fqs, stat_pe_fqs(raw),
// peWrite, stat_pe_fqs(clean) (peprocess.cpp:1915-1961) - become one engine submission.
#ifndef SNK_PROCESS_H
#define SNK_PROCESS_H
#include "cli_params.h"

namespace snk {

class FilterRun;      // shared implementation

class peProcess {
public:
    explicit peProcess(const HostParams& hp);
    ~peProcess();
    void process();
private:
    FilterRun* run_;
};

class seProcess {
public:
    explicit seProcess(const HostParams& hp);
    ~seProcess();
    void process();
private:
    FilterRun* run_;
};

}
#endif

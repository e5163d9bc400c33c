// gpu_process.h — the reference-side binding: subclasses of the REFERENCE's own peProcess / seProcess
// (/root/reference/src/peprocess.h:56, seprocess.h:32) whose per-batch work goes through the C ABI of
// include/snk_engine.h. Compiled against the reference sources by tests/integration/build.sh (which states the
// patch the reference needs: five `virtual` keywords and the two constructor calls in main.cpp) into oracle/_ref/SOAPnuke_gpu; tests/test_integration_gpu.py runs that
// binary against the unmodified reference binary. Test / integration material, not part of the product.
//
// What stays the reference's: argv parsing, the reader threads and their block partition, temp files, `cat`, the
// emission order, update_stat / print_stat (all ten reports are written by the reference's own code).
// What the engine replaces: filter_pe_fqs (peprocess.cpp:1424-1484), stat_pe_fqs raw + clean (:1076-1423) and the
// per-thread accumulators merge_stat reads (:1994-2005); the SE twins seprocess.cpp:871-917, :632-869, :1235.
#ifndef SNK_GPU_PROCESS_H
#define SNK_GPU_PROCESS_H
#include "peprocess.h"
#include "seprocess.h"
#include "snk_engine.h"

class gpuPeProcess : public peProcess {
public:
    explicit gpuPeProcess(C_global_parameter m_gp);
    ~gpuPeProcess();
    void filter_pe_fqs(PEcalOption* opt) override;                       // virtual in the reference (peprocess.h:61)
    void* stat_pe_fqs(PEstatOption opt, string dataType) override;       // needs `virtual` in peprocess.h:60
    void merge_stat() override;                                          // needs `virtual` in peprocess.h:69
private:
    struct Impl;
    Impl* d_;
};

class gpuSeProcess : public seProcess {
public:
    explicit gpuSeProcess(C_global_parameter m_gp);
    ~gpuSeProcess();
    void filter_se_fqs(SEcalOption opt) override;                        // needs `virtual` in seprocess.h:40
    void* stat_se_fqs(SEstatOption opt, string dataType) override;       // seprocess.h:38
    void merge_stat() override;                                          // seprocess.h:49
private:
    struct Impl;
    Impl* d_;
};
#endif

// tcgen05 (5th-gen tensor core) engine for the MLP GEMMs — placeholder until the UMMA kernel lands:
// reports "unsupported" so the fp32 FFMA engine in mlp.cu runs.
#include "common.cuh"
namespace wd {
struct GemmA; struct Epi;
int tc_gemm(WdModel*, int, const GemmA&, const float*, int, int, int, const Epi&, int, int) { return WD_EUNSUPPORTED; }
}

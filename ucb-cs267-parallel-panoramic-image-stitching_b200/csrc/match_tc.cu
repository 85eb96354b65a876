// match_tc.cu — K4 tensor-core matcher (placeholder until the tcgen05 kernel lands).
#include "common.cuh"
namespace pano {
bool match_tc_available() { return false; }
void match_tc_device(cudaStream_t, const DevDescriptors&, const DevDescriptors&, unsigned long long*) {}
}  // namespace pano

#include "common.cuh"
namespace dabgpu {
int msc_init_constants() { return 0; }
int ofdm_init_constants() { return 0; }
}

// synth_oracle.cpp — TEST INFRASTRUCTURE.  The seeded integer-only input generator (csrc/synth.h: the definition of the
// synthetic workload, shared verbatim by the host and device generators of the product) exported from the oracle library,
// so that the CPU baseline / `bench.py --impl reference` arm can make its inputs without loading the CUDA library.
#include <stddef.h>
#include <stdint.h>

#include "synth.h"

extern "C" {

void orbo_synth_image(uint32_t seed, int view, int cols, int rows, int max_disp, uint8_t* dst, size_t step)
{
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) dst[(size_t)y * step + x] = orbx_synth::image_pixel(seed, view, x, y, max_disp);
}

void orbo_synth_descriptors(uint32_t seed, int is_query, long long first_row, long long n_rows, long long ndb, int plant_every, uint8_t* dst)
{
    uint32_t* w = reinterpret_cast<uint32_t*>(dst);
    for (long long r = 0; r < n_rows; ++r)
        for (uint32_t k = 0; k < 8; ++k)
            w[r * 8 + k] = is_query ? orbx_synth::query_word(seed, (uint32_t)(first_row + r), k, (uint32_t)ndb, (uint32_t)plant_every)
                                    : orbx_synth::desc_word(seed, (uint32_t)(first_row + r), k);
}

}  // extern "C"

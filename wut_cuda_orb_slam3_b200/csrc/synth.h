// synth.h — integer-only, seeded synthetic inputs (images and descriptor sets), bit-identical on host and device.
//
// The datasets the reference was run on (EuRoC, KITTI) are not available offline, so every benchmark/parity
// input is generated from a counter-based hash.  Everything is 32-bit integer arithmetic so the host generator
// (tests, CPU baseline) and the device generator (bench, resident inputs) agree bit for bit.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ORBX_HD __host__ __device__ __forceinline__
#else
#define ORBX_HD inline
#endif

namespace orbx_synth {

ORBX_HD uint32_t mix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
    return h;
}

ORBX_HD uint32_t hash3(uint32_t seed, uint32_t a, uint32_t b, uint32_t c)
{
    return mix32(seed ^ mix32(a * 0x9E3779B1U + mix32(b * 0x85EBCA77U + mix32(c + 0xC2B2AE3DU))));
}

// Bilinear value noise on a lattice of `cell` pixels (cell is a power of two: shift = log2(cell)), 0..255.
ORBX_HD int value_noise(uint32_t seed, uint32_t octave, int x, int y, int shift)
{
    const int cell = 1 << shift;
    const int ix = x >> shift, iy = y >> shift;           // x,y >= 0
    const int fx = x & (cell - 1), fy = y & (cell - 1);
    const int v00 = hash3(seed, octave, (uint32_t)ix, (uint32_t)iy) & 255;
    const int v10 = hash3(seed, octave, (uint32_t)(ix + 1), (uint32_t)iy) & 255;
    const int v01 = hash3(seed, octave, (uint32_t)ix, (uint32_t)(iy + 1)) & 255;
    const int v11 = hash3(seed, octave, (uint32_t)(ix + 1), (uint32_t)(iy + 1)) & 255;
    const int top = v00 * (cell - fx) + v10 * fx;
    const int bot = v01 * (cell - fx) + v11 * fx;
    return (top * (cell - fy) + bot * fy) >> (2 * shift);
}

// One pixel of the synthetic "world" texture at integer world coordinates (wx, wy) >= 0.
//  * three octaves of value noise (64/16/4-px lattices) give smooth structure plus corner-like blobs,
//  * a blocky 8-px "checker" term adds genuine high-contrast corners,
//  * per-pixel noise whose amplitude varies slowly over the image (128-px lattice) so that some 35-px
//    cells are low-contrast (minThFAST retry path, empty cells) and others are dense.
ORBX_HD uint8_t world_pixel(uint32_t seed, int wx, int wy)
{
    const int n64 = value_noise(seed, 1u, wx, wy, 6);
    const int n16 = value_noise(seed, 2u, wx, wy, 4);
    const int n4 = value_noise(seed, 3u, wx, wy, 2);
    const int contrast = value_noise(seed, 4u, wx, wy, 7);             // 0..255, slow
    const int blk = (int)(hash3(seed, 5u, (uint32_t)(wx >> 3), (uint32_t)(wy >> 3)) & 255);
    int base = (n64 * 5 + n16 * 3) >> 3;                               // 0..255 smooth
    // detail terms are scaled by the local contrast (0..255)
    int detail = ((n4 - 128) * 4 + (blk - 128) * 3) >> 2;              // about -220..220
    int amp = contrast > 88 ? contrast - 88 : 0;                       // 0..167 ; ~30% of the area is flat
    int v = base + ((detail * amp) >> 8);
    const int namp = 1 + (amp >> 4);                                   // 1..10 pixel noise
    const int pn = (int)(hash3(seed, 6u, (uint32_t)wx, (uint32_t)wy) % (uint32_t)(2 * namp + 1)) - namp;
    v += pn;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// Per-row-band disparity (in pixels) used to derive the right image of a stereo pair: right(x,y) = world(x + d(y), y)
// so a feature at left column uL appears at right column uL - d.
ORBX_HD int row_band_disparity(uint32_t seed, int y, int max_disp)
{
    return 2 + (int)(hash3(seed, 7u, (uint32_t)(y / 48), 0u) % (uint32_t)(max_disp - 1));
}

// Image pixel: view = 0 (left / mono) or 1 (right view of the stereo pair with the same seed).
ORBX_HD uint8_t image_pixel(uint32_t seed, int view, int x, int y, int max_disp)
{
    const int margin = 64;   // keeps world coordinates positive for view 1
    const int d = view ? row_band_disparity(seed, y, max_disp) : 0;
    return world_pixel(seed, x + margin + d, y + margin);
}

// 256-bit descriptors: word w (0..7) of row `row`.  Rows listed as "planted" are copies of a database row with k
// bits flipped so that ratio tests pass and fail and distance ties occur.
ORBX_HD uint32_t desc_word(uint32_t seed, uint32_t row, uint32_t w)
{
    return hash3(seed, 11u, row, w);
}

// Query word: every `plant_every`-th query is database row (q * 2654435761 mod ndb) with k = (q/plant_every)%61 bits flipped.
ORBX_HD uint32_t query_word(uint32_t seed, uint32_t q, uint32_t w, uint32_t ndb, uint32_t plant_every)
{
    if (plant_every == 0 || (q % plant_every) != 0) return hash3(seed, 12u, q, w);
    const uint32_t src = (uint32_t)(((uint64_t)q * 2654435761ull) % ndb);
    uint32_t v = desc_word(seed, src, w);
    const uint32_t k = (q / plant_every) % 61u;
    for (uint32_t i = 0; i < k; ++i) {
        const uint32_t bit = hash3(seed, 13u, q, i) & 255u;   // may repeat: flips cancel, distance <= k
        if ((bit >> 5) == w) v ^= 1u << (bit & 31u);
    }
    return v;
}

}  // namespace orbx_synth

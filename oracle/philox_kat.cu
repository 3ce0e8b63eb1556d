// oracle/philox_kat.cu — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Evaluates NVIDIA cuRAND's own Philox4x32-10 (curand_philox4x32_x.h, shipped with the CUDA
// toolkit) ON THE HOST for a list of (counter, key) pairs and prints them as JSON, to pin the
// Philox restatements (oracle/stream_replay.c and tiny_mc_b200/csrc/philox.cuh) against an
// independent implementation.  Build + run:  nvcc -o /tmp/philox_kat oracle/philox_kat.cu && /tmp/philox_kat
#include <cstdint>
#include <cstdio>
#include <vector>
#include <vector_types.h>
#define QUALIFIERS static inline __host__ __device__
#include <curand_philox4x32_x.h>

int main()
{
    struct Case { uint32_t c[4], k[2]; };
    std::vector<Case> cases = {
        {{0, 0, 0, 0}, {0, 0}},
        {{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, {0xffffffffu, 0xffffffffu}},
        {{0x243f6a88u, 0x85a308d3u, 0x13198a2e, 0x03707344}, {0xa4093822u, 0x299f31d0u}},
    };
    uint32_t s = 12345u;
    for (int i = 0; i < 29; ++i) {
        Case c;
        for (auto& w : c.c) { s = s * 1664525u + 1013904223u; w = s; }
        for (auto& w : c.k) { s = s * 1664525u + 1013904223u; w = s; }
        cases.push_back(c);
    }
    printf("{\"source\": \"cuRAND curand_Philox4x32_10 (host build of curand_philox4x32_x.h, CUDA 12.9)\", \"rounds\": 10, \"cases\": [\n");
    for (size_t i = 0; i < cases.size(); ++i) {
        const Case& c = cases[i];
        uint4 r = curand_Philox4x32_10(make_uint4(c.c[0], c.c[1], c.c[2], c.c[3]), make_uint2(c.k[0], c.k[1]));
        printf(" {\"ctr\": [%u, %u, %u, %u], \"key\": [%u, %u], \"out\": [%u, %u, %u, %u]}%s\n", c.c[0], c.c[1], c.c[2],
               c.c[3], c.k[0], c.k[1], r.x, r.y, r.z, r.w, i + 1 < cases.size() ? "," : "");
    }
    printf("]}\n");
    return 0;
}

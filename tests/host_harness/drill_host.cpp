// Host build of the per-voxel arithmetic of csrc/drill.cuh (TEST HARNESS, compiled by tests/test_align_drill.py with g++):
// the same drill_accumulate / drill_finish the CUDA kernel calls, looped over the voxels on the CPU.
#include "../../geobo_b200/csrc/drill.cuh"

extern "C" void drill_host(const double* voxelpos, long N, const double* coord, const double* data, long ns, const double* voxsize,
                           long tile, double* out) {
    double *sx = new double[ns + 1], *sy = new double[ns + 1], *sz = new double[ns + 1];
    for (long s = 0; s < ns; ++s) { sx[s] = coord[3 * s]; sy[s] = coord[3 * s + 1]; sz[s] = coord[3 * s + 2]; }
    for (long v = 0; v < N; ++v) {
        DrillAcc acc;
        acc.sum = 0.0;
        acc.cnt = 0;
        for (long t0 = 0; t0 < ns; t0 += tile) {               // tile by tile like the kernel's shared-memory staging
            const long n = ns - t0 < tile ? ns - t0 : tile;
            drill_accumulate(acc, voxelpos[v], voxelpos[N + v], voxelpos[2 * N + v], sx + t0, sy + t0, sz + t0, data + t0, n, voxsize[0],
                             voxsize[1], voxsize[2]);
        }
        out[v] = drill_finish(acc);
    }
    delete[] sx;
    delete[] sy;
    delete[] sz;
}

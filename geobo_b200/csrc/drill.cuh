// Per-voxel arithmetic of align_drill (geobo/run_geobo.py:132-159, geobo/utils.py:55-83), shared by the device kernel
// (drill.cu) and by the host harness the CPU tests compile (tests/host_harness/drill_host.cpp), so the lines that decide
// which samples a voxel owns are exercised without a GPU as well.
#pragma once
#if defined(__CUDACC__)
#define GB_HD __host__ __device__ __forceinline__
#else
#define GB_HD inline
#endif

// Running state of one voxel: sum and count of the non-NaN samples inside its window (np.nanmean = sum / count).
struct DrillAcc {
    double sum;
    long cnt;
};

// Adds the samples [0, n) of one tile.  A sample belongs to the voxel centred at (vx, vy, vz) when, per axis,
// centre - d <= coordinate < centre + d  (run_geobo.py:149-151: a window of TWO voxel sizes, half open).
GB_HD void drill_accumulate(DrillAcc& acc, double vx, double vy, double vz, const double* cx, const double* cy, const double* cz,
                            const double* val, long n, double dx, double dy, double dz) {
    const double x0 = vx - dx, x1 = vx + dx, y0 = vy - dy, y1 = vy + dy, z0 = vz - dz, z1 = vz + dz;
    for (long s = 0; s < n; ++s) {
        const double x = cx[s], y = cy[s], z = cz[s];
        if (x0 <= x && x < x1 && y0 <= y && y < y1 && z0 <= z && z < z1) {
            const double v = val[s];
            if (v == v) {          // nanmean skips NaN
                acc.sum += v;
                acc.cnt += 1;
            }
        }
    }
}

// run_geobo.py:152-156: the voxel keeps 0 unless the mean exists and is finite (no sample, only NaN samples or an infinite
// mean leave the zero).
GB_HD double drill_finish(const DrillAcc& acc) {
    if (acc.cnt == 0) return 0.0;
    const double m = acc.sum / (double)acc.cnt;
    return (m - m == 0.0) ? m : 0.0;
}

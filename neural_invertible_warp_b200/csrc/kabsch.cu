// Batched rigid registration (the global-alignment fit that replaces roma).
//   reference: roma.rigid_points_registration at model/nerf_inn_llff.py:569 and model/pose_models/inn.py:100
//              (roma==1.4.1: centroids, cross-covariance, SVD, det fix), results detached.
// The least-squares proper rotation R and translation t with y ~ R x + t per image.  One block per image: two
// reduction passes (centroids, centred cross-covariance) and Horn's closed form -- R is the rotation of the unit
// quaternion that is the dominant eigenvector of a symmetric 4x4 built from the cross-covariance -- solved by cyclic
// Jacobi in fp64 by one thread.  The optimum is the same as Kabsch's U diag(1,1,det) V^T; no SVD library call, no host
// synchronisation, capturable in a CUDA graph.
#include "common.cuh"

namespace {

__device__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;
}

// Horn (1987): the rotation x -> y maximising sum y . R x is the dominant eigenvector of a symmetric 4x4 built from the
// cross-covariance S[a][c] = sum (x_a - mx_a)(y_c - my_c); cyclic Jacobi in fp64 by one thread
__device__ void horn_solve(const double (&S)[3][3], const double (&mx)[3], const double (&my)[3], int b,
                           float* __restrict__ R, float* __restrict__ t) {
    // Horn (1987): the rotation x -> y maximising sum y . R x is the dominant eigenvector of N
    double N[4][4] = {
        {S[0][0] + S[1][1] + S[2][2], S[1][2] - S[2][1], S[2][0] - S[0][2], S[0][1] - S[1][0]},
        {0, S[0][0] - S[1][1] - S[2][2], S[0][1] + S[1][0], S[2][0] + S[0][2]},
        {0, 0, -S[0][0] + S[1][1] - S[2][2], S[1][2] + S[2][1]},
        {0, 0, 0, -S[0][0] - S[1][1] + S[2][2]}};
    for (int i = 0; i < 4; ++i) for (int j = 0; j < i; ++j) N[i][j] = N[j][i];
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 4; ++p) for (int q = p + 1; q < 4; ++q) off += N[p][q] * N[p][q];
        if (off < 1e-30) break;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) {
                if (fabs(N[p][q]) < 1e-300) continue;
                const double theta = (N[q][q] - N[p][p]) / (2.0 * N[p][q]);
                const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(tt * tt + 1.0), sn = tt * cs;
                for (int k = 0; k < 4; ++k) {                     // N <- J^T N J
                    const double a = N[k][p], c = N[k][q];
                    N[k][p] = cs * a - sn * c; N[k][q] = sn * a + cs * c;
                }
                for (int k = 0; k < 4; ++k) {
                    const double a = N[p][k], c = N[q][k];
                    N[p][k] = cs * a - sn * c; N[q][k] = sn * a + cs * c;
                }
                for (int k = 0; k < 4; ++k) {
                    const double a = V[k][p], c = V[k][q];
                    V[k][p] = cs * a - sn * c; V[k][q] = sn * a + cs * c;
                }
            }
    }
    int best = 0;
    for (int i = 1; i < 4; ++i) if (N[i][i] > N[best][best]) best = i;
    double q0 = V[0][best], q1 = V[1][best], q2 = V[2][best], q3 = V[3][best];
    const double n = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    q0 /= n; q1 /= n; q2 /= n; q3 /= n;
    const double Rm[3][3] = {
        {q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)},
        {2 * (q1 * q2 + q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * (q2 * q3 - q0 * q1)},
        {2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3}};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[b * 9 + i * 3 + j] = (float)Rm[i][j];
        t[b * 3 + i] = (float)(my[i] - (Rm[i][0] * mx[0] + Rm[i][1] * mx[1] + Rm[i][2] * mx[2]));
    }
}

__global__ void __launch_bounds__(128)
kabsch_kernel(const float* __restrict__ x, const float* __restrict__ y, int M, float* __restrict__ R, float* __restrict__ t) {
    __shared__ double red[4];
    const int b = blockIdx.x;
    const float* xb = x + (size_t)b * M * 3;
    const float* yb = y + (size_t)b * M * 3;
    double mx[3], my[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double sx = 0.0, sy = 0.0;
        for (int i = threadIdx.x; i < M; i += blockDim.x) { sx += xb[i * 3 + c]; sy += yb[i * 3 + c]; }
        mx[c] = block_sum(sx, red) / M;
        my[c] = block_sum(sy, red) / M;
    }
    double S[3][3];                                   // S[a][b] = sum (x_a - mx_a)(y_b - my_b)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double s = 0.0;
            for (int i = threadIdx.x; i < M; i += blockDim.x) s += ((double)xb[i * 3 + a] - mx[a]) * ((double)yb[i * 3 + c] - my[c]);
            S[a][c] = block_sum(s, red);
        }
    if (threadIdx.x != 0) return;
    horn_solve(S, mx, my, b, R, t);
}

// ---- the same fit from per-image sufficient statistics, so that a point list sharded over ranks (data-parallel ray
// shards) is fitted EXACTLY as the whole list: every rank accumulates stats[b] = (n, sum x [3], sum y [3], sum x_a y_c [9]) in
// fp64 over its rows, the 16 doubles per image are summed over the ranks (one all-reduce), and every rank solves
// S[a][c] = sum x_a y_c - (sum x_a)(sum y_c) / n.  
__global__ void __launch_bounds__(128)
kabsch_stats_kernel(const float* __restrict__ x, const float* __restrict__ y, int M, double* __restrict__ stats) {
    __shared__ double red[4];
    const int b = blockIdx.x;
    const float* xb = x + (size_t)b * M * 3;
    const float* yb = y + (size_t)b * M * 3;
    double acc[15];
#pragma unroll
    for (int k = 0; k < 15; ++k) acc[k] = 0.0;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        const double xv[3] = {xb[i * 3], xb[i * 3 + 1], xb[i * 3 + 2]}, yv[3] = {yb[i * 3], yb[i * 3 + 1], yb[i * 3 + 2]};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            acc[a] += xv[a]; acc[3 + a] += yv[a];
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[6 + a * 3 + c] += xv[a] * yv[c];
        }
    }
#pragma unroll
    for (int k = 0; k < 15; ++k) {
        const double v = block_sum(acc[k], red);
        if (threadIdx.x == 0) stats[b * 16 + 1 + k] = v;
    }
    if (threadIdx.x == 0) stats[b * 16] = (double)M;
}

__global__ void kabsch_solve_kernel(const double* __restrict__ stats, int B, float* __restrict__ R, float* __restrict__ t) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* s = stats + b * 16;
    const double n = s[0];
    double mx[3], my[3], S[3][3];
    for (int a = 0; a < 3; ++a) { mx[a] = s[1 + a] / n; my[a] = s[4 + a] / n; }
    for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 3; ++c) S[a][c] = s[7 + a * 3 + c] - n * mx[a] * my[c];
    horn_solve(S, mx, my, b, R, t);
}

}  // namespace

extern "C" int niw_kabsch_stats(const float* x, const float* y, int B, int M, double* stats, void* stream) {
    NIW_CHECK_ARG(x && y && stats && B > 0 && M > 0);
    niw::note_launch(), kabsch_stats_kernel<<<B, 128, 0, niw_stream(stream)>>>(x, y, M, stats);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_kabsch_solve(const double* stats, int B, float* R, float* t, void* stream) {
    NIW_CHECK_ARG(stats && R && t && B > 0);
    niw::note_launch(), kabsch_solve_kernel<<<niw_blocks(B, 32), 32, 0, niw_stream(stream)>>>(stats, B, R, t);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_kabsch(const float* x, const float* y, int B, int M, float* R, float* t, void* stream) {
    NIW_CHECK_ARG(x && y && R && t && B > 0 && M > 0);
    niw::note_launch(), kabsch_kernel<<<B, 128, 0, niw_stream(stream)>>>(x, y, M, R, t);
    NIW_LAUNCH_CHECK();
    return 0;
}

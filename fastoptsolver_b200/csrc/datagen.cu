// datagen.cu -- synthetic correlated-column designs generated directly in HBM, and the
// repack kernel used when a host matrix arrives in Fortran order.
//
// The distributional recipe is the reference's 5-column generator (easy_boston_data.py:23-43)
// generalised to d columns and population-standardised (zero mean, unit variance per
// column, which is what the missing notebook did to its data):
//   per group of five columns:  (z1, r1 z1 + sqrt(1-r1^2) z2), (z3, r2 z3 + sqrt(1-r2^2) z4), z5
//   x_true = tile([5, 0, -0.02, -0.05, 1.5]);  b = A x_true + noise_std * N(0,1)
// Randomness: Philox4x32-10 keyed by the seed, counter = (global row, group, draw), so any
// row range of the virtual matrix can be produced independently (row-sharded ranks).
#include <string.h>

#include "fos_common.cuh"

namespace {

struct U4 {
    unsigned x, y, z, w;
};

__device__ __forceinline__ U4 philox4x32_10(U4 ctr, unsigned k0, unsigned k1) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const unsigned W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        U4 n;
        n.x = hi1 ^ ctr.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ ctr.w ^ k1;
        n.w = lo0;
        ctr = n;
        k0 += W0;
        k1 += W1;
    }
    return ctr;
}

// two standard normals from two 32-bit words (Box-Muller in double precision)
__device__ __forceinline__ void normal2(unsigned a, unsigned b, double& n0, double& n1) {
    const double u1 = (static_cast<double>(a) + 1.0) * (1.0 / 4294967296.0);  // (0, 1]
    const double u2 = static_cast<double>(b) * (1.0 / 4294967296.0);          // [0, 1)
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    n0 = rad * c;
    n1 = rad * s;
}

template <typename T>
__global__ void synth_kernel(T* __restrict__ A, double* __restrict__ b, long long n, int d, int lda,
                             unsigned long long seed, double noise, double rho1, double rho2,
                             long long row0) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const unsigned k0 = static_cast<unsigned>(seed), k1 = static_cast<unsigned>(seed >> 32);
    const double c1 = sqrt(1.0 - rho1 * rho1), c2 = sqrt(1.0 - rho2 * rho2);
    const int groups = d / 5, rest = d - 5 * groups;
    const double coef[5] = {5.0, 0.0, -0.02, -0.05, 1.5};

    for (long long row = warp; row < n; row += nwarps) {
        const unsigned long long grow = static_cast<unsigned long long>(row0 + row);
        T* out = A + row * lda;
        double dot = 0.0;
        for (int g = lane; g < groups; g += 32) {
            U4 c;
            c.x = static_cast<unsigned>(grow);
            c.y = static_cast<unsigned>(grow >> 32);
            c.z = static_cast<unsigned>(g);
            c.w = 0u;
            const U4 ra = philox4x32_10(c, k0, k1);
            c.w = 1u;
            const U4 rb = philox4x32_10(c, k0, k1);
            double z[6];
            normal2(ra.x, ra.y, z[0], z[1]);
            normal2(ra.z, ra.w, z[2], z[3]);
            normal2(rb.x, rb.y, z[4], z[5]);
            double col[5];
            col[0] = z[0];
            col[1] = rho1 * z[0] + c1 * z[1];
            col[2] = z[2];
            col[3] = rho2 * z[2] + c2 * z[3];
            col[4] = z[4];
#pragma unroll
            for (int e = 0; e < 5; ++e) {
                const T stored = static_cast<T>(col[e]);
                out[5 * g + e] = stored;
                dot = fma(static_cast<double>(stored), coef[e], dot);
            }
        }
        for (int j = lane; j < rest; j += 32) {  // left-over columns: N(0,1), zero coefficient
            U4 c;
            c.x = static_cast<unsigned>(grow);
            c.y = static_cast<unsigned>(grow >> 32);
            c.z = static_cast<unsigned>(groups + j);
            c.w = 2u;
            const U4 ra = philox4x32_10(c, k0, k1);
            double z0, z1;
            normal2(ra.x, ra.y, z0, z1);
            out[5 * groups + j] = static_cast<T>(z0);
        }
        for (int j = d + lane; j < lda; j += 32) out[j] = static_cast<T>(0);
        dot = fos_warp_sum(dot);
        if (lane == 0) {
            U4 c;
            c.x = static_cast<unsigned>(grow);
            c.y = static_cast<unsigned>(grow >> 32);
            c.z = 0xFFFFFFFFu;
            c.w = 3u;
            const U4 ra = philox4x32_10(c, k0, k1);
            double z0, z1;
            normal2(ra.x, ra.y, z0, z1);
            b[row] = dot + noise * z0;
        }
    }
}

// src: column-major block [d][rows] (leading dimension = rows); dst: row-major [rows][lda]
template <typename T>
__global__ void repack_cm_kernel(const T* __restrict__ src, T* __restrict__ dst, long long rows, int d,
                                 int lda) {
    __shared__ T tile[32][33];
    const long long r0 = static_cast<long long>(blockIdx.x) * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long r = r0 + threadIdx.x;
        const int c = c0 + j;
        tile[j][threadIdx.x] = (r < rows && c < d) ? src[static_cast<size_t>(c) * rows + r] : static_cast<T>(0);
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long r = r0 + j;
        const int c = c0 + threadIdx.x;
        if (r < rows && c < lda) dst[static_cast<size_t>(r) * lda + c] = (c < d) ? tile[threadIdx.x][j] : static_cast<T>(0);
    }
}

// ---------------------------------------------------------------- column statistics / z-scoring
// part[cta][c] = sum over the CTA's row block of (A[i][c] - center[c])^p, p = 1 or 2.  Fixed
// row partition and fixed-order final sum: deterministic.  One pass over A (HBM bound).
template <typename T>
__global__ void colstat_kernel(const T* __restrict__ A, long long n, int d, int lda, const double* __restrict__ center,
                               int squared, double* __restrict__ part) {
    const long long lo = (n * blockIdx.x) / gridDim.x, hi = (n * (blockIdx.x + 1LL)) / gridDim.x;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const double mu = center ? center[c] : 0.0;
        double s0 = 0.0, s1 = 0.0;
        long long r = lo;
        for (; r + 1 < hi; r += 2) {
            const double v0 = static_cast<double>(A[r * lda + c]) - mu;
            const double v1 = static_cast<double>(A[(r + 1) * lda + c]) - mu;
            s0 += squared ? v0 * v0 : v0;
            s1 += squared ? v1 * v1 : v1;
        }
        if (r < hi) {
            const double v0 = static_cast<double>(A[r * lda + c]) - mu;
            s0 += squared ? v0 * v0 : v0;
        }
        part[static_cast<size_t>(blockIdx.x) * d + c] = s0 + s1;
    }
}

__global__ void colstat_reduce_kernel(const double* __restrict__ part, int nparts, int d, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += part[static_cast<size_t>(p) * d + c];
    out[c] = s;
}

// A[i][c] = (A[i][c] - shift[c]) / scale[c] in place (numpy: (A - mu) / sd)
template <typename T>
__global__ void affine_columns_kernel(T* __restrict__ A, long long n, int d, int lda, const double* __restrict__ shift,
                                      const double* __restrict__ scale) {
    const long long lo = (n * blockIdx.x) / gridDim.x, hi = (n * (blockIdx.x + 1LL)) / gridDim.x;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const double mu = shift[c], sd = scale[c];
        for (long long r = lo; r < hi; ++r) {
            const double v = static_cast<double>(A[r * lda + c]);
            A[r * lda + c] = static_cast<T>(__ddiv_rn(__dsub_rn(v, mu), sd));
        }
    }
}

}  // namespace

int fos_launch_synthetic(fos_design* h, unsigned long long seed, double noise, double rho1, double rho2,
                         long long row0) {
    const int threads = 256;
    long long want = (h->n + 7) / 8;
    long long cap = static_cast<long long>(h->sm_count) * 16;
    const unsigned blocks = static_cast<unsigned>(want < 1 ? 1 : (want > cap ? cap : want));
    if (h->dtype == FOS_F64)
        synth_kernel<double><<<blocks, threads, 0, h->stream>>>(static_cast<double*>(h->A), h->b, h->n, h->d, h->lda,
                                                                seed, noise, rho1, rho2, row0);
    else
        synth_kernel<float><<<blocks, threads, 0, h->stream>>>(static_cast<float*>(h->A), h->b, h->n, h->d, h->lda,
                                                               seed, noise, rho1, rho2, row0);
    FOS_CUDA(cudaGetLastError());
    return FOS_OK;
}

int fos_launch_repack(const void* src_dev, void* dst_dev, long long rows, int d, int lda, long long, long long,
                      int dtype, cudaStream_t s) {
    dim3 grid(static_cast<unsigned>((rows + 31) / 32), static_cast<unsigned>((lda + 31) / 32));
    dim3 block(32, 8);
    if (dtype == FOS_F64)
        repack_cm_kernel<double><<<grid, block, 0, s>>>(static_cast<const double*>(src_dev),
                                                        static_cast<double*>(dst_dev), rows, d, lda);
    else
        repack_cm_kernel<float><<<grid, block, 0, s>>>(static_cast<const float*>(src_dev),
                                                       static_cast<float*>(dst_dev), rows, d, lda);
    FOS_CUDA(cudaGetLastError());
    return FOS_OK;
}

extern "C" int fos_design_column_sums(fos_design* h, const double* center, int squared, double* out, double* b_out) {
    FOS_REQUIRE(h && out, "null pointer argument");
    FOS_CUDA(cudaSetDevice(h->device));
    const int d = h->d;
    const int grid = h->sm_count * 4;
    double *part = nullptr, *dout = nullptr, *dcen = nullptr;
    auto cleanup = [&]() {
        for (double* q : {part, dout, dcen})
            if (q) cudaFree(q);
    };
    auto body = [&]() -> int {
        const int dd = d + 1;  // last slot: the same statistic of b
        FOS_CUDA(cudaMalloc(&part, static_cast<size_t>(grid) * dd * sizeof(double)));
        FOS_CUDA(cudaMalloc(&dout, dd * sizeof(double)));
        if (center) {
            FOS_CUDA(cudaMalloc(&dcen, dd * sizeof(double)));
            FOS_CUDA(cudaMemcpyAsync(dcen, center, dd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        }
        if (h->dtype == FOS_F64)
            colstat_kernel<double><<<grid, 256, 0, h->stream>>>(static_cast<const double*>(h->A), h->n, d, h->lda, dcen,
                                                              squared, part);
        else
            colstat_kernel<float><<<grid, 256, 0, h->stream>>>(static_cast<const float*>(h->A), h->n, d, h->lda, dcen,
                                                             squared, part);
        colstat_reduce_kernel<<<(d + 255) / 256, 256, 0, h->stream>>>(part, grid, d, dout);
        // b as an n x 1 matrix
        colstat_kernel<double><<<grid, 32, 0, h->stream>>>(h->b, h->n, 1, 1, dcen ? dcen + d : nullptr, squared,
                                                          part + static_cast<size_t>(grid) * d);
        colstat_reduce_kernel<<<1, 32, 0, h->stream>>>(part + static_cast<size_t>(grid) * d, grid, 1, dout + d);
        FOS_CUDA(cudaGetLastError());
        std::vector<double> host(dd);
        FOS_CUDA(cudaMemcpyAsync(host.data(), dout, dd * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        FOS_CUDA(cudaStreamSynchronize(h->stream));
        memcpy(out, host.data(), d * sizeof(double));
        if (b_out) *b_out = host[d];
        return FOS_OK;
    };
    const int st = body();
    cleanup();
    return st;
}

extern "C" int fos_design_affine(fos_design* h, const double* shift, const double* scale, double b_shift) {
    FOS_REQUIRE(h && shift && scale, "null pointer argument");
    FOS_REQUIRE(h->owns_A, "cannot rewrite a borrowed matrix in place");
    FOS_CUDA(cudaSetDevice(h->device));
    fos_upload_gram_drop(h);  // A changes: a Gram matrix accumulated under the upload is stale
    const int d = h->d;
    double* dv = nullptr;
    FOS_CUDA(cudaMalloc(&dv, (2 * static_cast<size_t>(d) + 2) * sizeof(double)));
    auto body = [&]() -> int {
        FOS_CUDA(cudaMemcpyAsync(dv, shift, d * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        FOS_CUDA(cudaMemcpyAsync(dv + d, scale, d * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        const double bs[2] = {b_shift, 1.0};
        FOS_CUDA(cudaMemcpyAsync(dv + 2 * d, bs, 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        const int grid = h->sm_count * 4;
        if (h->dtype == FOS_F64)
            affine_columns_kernel<double><<<grid, 256, 0, h->stream>>>(static_cast<double*>(h->A), h->n, d, h->lda, dv,
                                                                     dv + d);
        else
            affine_columns_kernel<float><<<grid, 256, 0, h->stream>>>(static_cast<float*>(h->A), h->n, d, h->lda, dv,
                                                                    dv + d);
        affine_columns_kernel<double><<<grid, 32, 0, h->stream>>>(h->b, h->n, 1, 1, dv + 2 * d, dv + 2 * d + 1);
        FOS_CUDA(cudaGetLastError());
        FOS_CUDA(cudaStreamSynchronize(h->stream));
        return FOS_OK;
    };
    const int st = body();
    cudaFree(dv);
    return st;
}

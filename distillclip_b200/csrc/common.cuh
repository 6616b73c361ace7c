// Shared device/host helpers for the distillclip_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/distillclip_b200.h"

namespace dcb {

// ------------------------------------------------------------------------------------------
// error reporting (thread-local message behind dcb_last_error())
// ------------------------------------------------------------------------------------------
char* error_buffer();   // capi.cu
inline int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return 1;
}
#define DCB_CUDA_OK(expr)                                                                        \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess) return ::dcb::fail("%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                                 __FILE__, __LINE__);                            \
    } while (0)
#define DCB_REQUIRE(cond, ...)                    \
    do {                                          \
        if (!(cond)) return ::dcb::fail(__VA_ARGS__); \
    } while (0)

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

inline int dtype_size(int dt) { return dt == DCB_F32 ? 4 : 2; }

// ------------------------------------------------------------------------------------------
// element conversion
// ------------------------------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<__nv_bfloat16> {
    static constexpr int kPer16B = 8;
    __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Elem<__half> {
    static constexpr int kPer16B = 8;
    __device__ static __forceinline__ float to_f(__half v) { return __half2float(v); }
    __device__ static __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct Elem<float> {
    static constexpr int kPer16B = 4;
    __device__ static __forceinline__ float to_f(float v) { return v; }
    __device__ static __forceinline__ float from_f(float v) { return v; }
};

// Unpack a 32-bit word holding two 16-bit floats.
template <typename T> __device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi);
template <> __device__ __forceinline__ void unpack2<__nv_bfloat16>(uint32_t w, float& lo, float& hi) {
    lo = __uint_as_float(w << 16);
    hi = __uint_as_float(w & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack2<__half>(uint32_t w, float& lo, float& hi) {
    float2 f = __half22float2(*reinterpret_cast<__half2*>(&w));
    lo = f.x;
    hi = f.y;
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// N consecutive elements of type T at `p` (aligned to N*sizeof(T)) -> floats.  N*sizeof(T) in {2,4,8,16}.
template <typename T, int N> __device__ __forceinline__ void load_vec(const T* p, float (&out)[N]) {
    constexpr int kBytes = N * (int)sizeof(T);
    if constexpr (kBytes == 16) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        if constexpr (sizeof(T) == 4) {
            out[0] = __uint_as_float(v.x); out[1] = __uint_as_float(v.y);
            out[2] = __uint_as_float(v.z); out[3] = __uint_as_float(v.w);
        } else {
            unpack2<T>(v.x, out[0], out[1]); unpack2<T>(v.y, out[2], out[3]);
            unpack2<T>(v.z, out[4], out[5]); unpack2<T>(v.w, out[6], out[7]);
        }
    } else if constexpr (kBytes == 8) {
        uint2 v;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
        if constexpr (sizeof(T) == 4) {
            out[0] = __uint_as_float(v.x); out[1] = __uint_as_float(v.y);
        } else {
            unpack2<T>(v.x, out[0], out[1]); unpack2<T>(v.y, out[2], out[3]);
        }
    } else if constexpr (kBytes == 4) {
        uint32_t v;
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
        if constexpr (sizeof(T) == 4) out[0] = __uint_as_float(v);
        else unpack2<T>(v, out[0], out[1]);
    } else {
        static_assert(kBytes == 2, "unsupported vector width");
        out[0] = Elem<T>::to_f(__ldg(p));
    }
}

// Store N floats as N consecutive elements of type G at `p` (aligned to min(16, N*sizeof(G))).
template <typename G, int N> __device__ __forceinline__ void store_vec(G* p, const float (&v)[N]) {
    constexpr int kBytes = N * (int)sizeof(G);
    if constexpr (sizeof(G) == 4) {
        if constexpr (N % 4 == 0) {
#pragma unroll
            for (int i = 0; i < N; i += 4)
                *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else if constexpr (N % 2 == 0) {
#pragma unroll
            for (int i = 0; i < N; i += 2) *reinterpret_cast<float2*>(p + i) = make_float2(v[i], v[i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) p[i] = v[i];
        }
    } else {
        if constexpr (kBytes == 16) {
            uint4 w = make_uint4(pack2<G>(v[0], v[1]), pack2<G>(v[2], v[3]), pack2<G>(v[4], v[5]), pack2<G>(v[6], v[7]));
            *reinterpret_cast<uint4*>(p) = w;
        } else if constexpr (kBytes == 8) {
            *reinterpret_cast<uint2*>(p) = make_uint2(pack2<G>(v[0], v[1]), pack2<G>(v[2], v[3]));
        } else if constexpr (kBytes == 4) {
            *reinterpret_cast<uint32_t*>(p) = pack2<G>(v[0], v[1]);
        } else {
            static_assert(N == 1, "unsupported vector width");
            p[0] = Elem<G>::from_f(v[0]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over a block (blockDim.x multiple of 32, <= 1024). Result valid in thread 0.
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double warp_part[32];
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_part[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? warp_part[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// dtype dispatch: calls f(T{}, G{}) with the element types for (in_dtype, grad_dtype).
template <typename F> inline int dispatch_in_grad(int in_dtype, int grad_dtype, F&& f) {
    if (grad_dtype != in_dtype && grad_dtype != DCB_F32) return fail("grad_dtype must equal in_dtype or be DCB_F32");
    switch (in_dtype) {
        case DCB_BF16:
            return grad_dtype == DCB_F32 ? f(__nv_bfloat16{}, float{}) : f(__nv_bfloat16{}, __nv_bfloat16{});
        case DCB_F16:
            return grad_dtype == DCB_F32 ? f(__half{}, float{}) : f(__half{}, __half{});
        case DCB_F32:
            return f(float{}, float{});
        default:
            return fail("unknown dtype %d", in_dtype);
    }
}

}  // namespace dcb

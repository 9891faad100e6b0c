// k_misc.cu -- primary rays, final shading, ray queries and parity dump kernels.
#include "reservoir.cuh"
#include "launch.hpp"

namespace romis {

// ------------------------------------------------------------------------------------------------
// primary rays -> G-buffer (band rows plus halo rows)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) primary_kernel(SceneDev sc, FrameDev fr, GBufDev g, int row0, int row1) {
    int x, y; thread_pixel<false>(x, y);
    y += row0;
    if (x >= fr.W || y >= row1) return;
    v3 d = gen_ray_dir(fr.cam, x, y, fr.W, fr.H);
    float t, u, v; uint32_t tri;
    size_t p = (size_t)(y - fr.ey0) * fr.W + x;
    const bool hit = trace_closest(sc, fr.cam.origin, d, FLT_MAX, t, u, v, tri);
    pdl_wait();                                     // the kernel before this one (the previous frame's) may still read the G-buffer
    pdl_launch_dependents();
    if (hit) {
        const float4* a = sc.tri_attr + 4 * (size_t)tri;
        float4 a0 = __ldg(a), a1 = __ldg(a + 1), a2 = __ldg(a + 2), a3 = __ldg(a + 3);
        float w = (1.0f - u) - v;                       // attribute interpolation (w*a + u*b) + v*c, oracle/tracer.h
        v3 n = add3(add3(scale3(V3(a0.x, a0.y, a0.z), w), scale3(V3(a1.x, a1.y, a1.z), u)), scale3(V3(a2.x, a2.y, a2.z), v));
        g.tn[p] = make_float4(t, n.x, n.y, n.z);
        g.mesh[p] = __float_as_uint(a3.w);
        if (sc.has_textures) {
            float tu = (w * a0.w + u * a2.w) + v * a3.y;
            float tv = (w * a1.w + u * a3.x) + v * a3.z;
            g.uv[p] = make_float2(tu, tv);
        }
    } else {                                            // miss: value-initialised RayHit
        g.tn[p] = make_float4(FLT_MAX, 0.0f, 0.0f, 0.0f);
        g.mesh[p] = (uint32_t)sc.n_meshes;
        if (sc.has_textures) g.uv[p] = make_float2(0.0f, 0.0f);
    }
}

// ------------------------------------------------------------------------------------------------
// final shading + tone mapping -> Screen layout (row-flipped float RGB)
// ------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(ROMIS_LBT_SHADE, ROMIS_MINB_SHADE) shade_kernel(SceneDev sc, FrameDev fr, GBufDev g, ResBuf in, float* __restrict__ rgb, FineDev fd) {
    int x, y; thread_pixel<false>(x, y);
    y += fr.y0;
    if (x >= fr.W || y >= fr.y1) return;
    const int N = NT > 0 ? NT : (int)fr.f.numSamplesInReservoir;
    const bool es = fr.f.enableShading != 0;
    const int lrow = y - fr.ey0;
    PixCtx c = make_ctx(sc, fr, g, x, y);            // G-buffer only: older than the previous kernel
    { const int by0 = fr.y0 + (int)(blockIdx.y * blockDim.y); fine_wait(fd, by0, by0 + (int)blockDim.y); }
    pdl_launch_dependents();
    v3 color = V3(0, 0, 0);
    // Miss pixels shade to exactly +0 (computeShading is 0, W is 0); a sample with W == 0 adds (+-0) and leaves the sum
    // unchanged whatever its visibility: both skip the shadow ray and the shading without changing a bit.
    _Pragma("unroll 1") for (int j = 0; j < N && !c.miss; j++) {                   // render_utils.cpp:56-62 (rolled: one copy of the shading code)
        uint4 rec = res_rec(in, lrow, j)[x];
        if (__uint_as_float(rec.w) == 0.0f) continue;
        v3 pos, col; light_sample(sc, rec.x, __uint_as_float(rec.y), __uint_as_float(rec.z), pos, col);
        v3 scol = visible(sc, c, pos) ? compute_shading(c, es, pos, col) : V3(0, 0, 0);
        scol = scale3(scol, __uint_as_float(rec.w));
        color = add3(color, scol);
    }
    color = div3(color, (float)N);                                                  // :63
    if (fr.f.enableToneMapping) color = tone_map(color, fr.f);                      // tone_mapping.cpp:8-11
    size_t i = (size_t)(fr.H - 1 - y) * fr.W + x;                                   // screen.cpp:37-43
    rgb[3 * i] = color.x; rgb[3 * i + 1] = color.y; rgb[3 * i + 2] = color.z;
}

// ray queries for the tracer parity tests (romis_trace_rays)
__global__ void trace_kernel(SceneDev sc, const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ tfar, int n,
                             int any_hit, uint8_t* hit, float* t, float* u, float* v, uint32_t* tri) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    v3 oo = V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), dd = V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
    if (any_hit) { hit[i] = trace_any(sc, oo, dd, tfar[i]) ? 1 : 0; return; }
    float tt, uu, vv; uint32_t ti;
    bool h = trace_closest(sc, oo, dd, tfar[i], tt, uu, vv, ti);
    hit[i] = h ? 1 : 0;
    if (h) { t[i] = tt; u[i] = uu; v[i] = vv; tri[i] = ti; }
}

// div3_shared against the plain division (romis_selftest_division): out_fast / out_ref get the bits of both routes
__global__ void division_selftest_kernel(const float* __restrict__ num, const float* __restrict__ den, int n, float* out_fast, float* out_ref) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const v3 a = V3(num[3 * i], num[3 * i + 1], num[3 * i + 2]);
    const float d = den[i];
    const v3 f = div3_shared(a, d), r = div3(a, d);
    out_fast[3 * i] = f.x; out_fast[3 * i + 1] = f.y; out_fast[3 * i + 2] = f.z;
    out_ref[3 * i] = r.x; out_ref[3 * i + 1] = r.y; out_ref[3 * i + 2] = r.z;
}
void launch_division_selftest(cudaStream_t s, const float* num, const float* den, int n, float* out_fast, float* out_ref) {
    division_selftest_kernel<<<(n + 255) / 256, 256, 0, s>>>(num, den, n, out_fast, out_ref);
}

// hit pixels per image row (romis_row_hit_counts): one warp per row
__global__ void row_hits_kernel(GBufDev g, int W, int H, uint32_t n_meshes, uint32_t* rows) {
    int y = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (y >= H) return;
    uint32_t n = 0;
    for (int x = lane; x < W; x += 32) n += g.mesh[(size_t)y * W + x] != n_meshes;
    for (int o = 16; o; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
    if (lane == 0) rows[y] = n;
}

// Flag words of the peer-mapped halo protocol (romis_gpu.cu peer_exchange).  signal: publish a token into a neighbour's
// flag word after everything earlier in the stream (the halo push) has completed.  wait: hold the stream until both
// neighbours' tokens have arrived; the writer runs on ANOTHER GPU, so the spin cannot starve it.  A bounded spin
// (~2 s) records a timeout instead of hanging the device if a neighbour never arrives.
__global__ void signal_kernel(volatile uint32_t* a, uint32_t va, volatile uint32_t* b, uint32_t vb) {
    __threadfence_system();
    if (a) *a = va;
    if (b) *b = vb;
    __threadfence_system();
}
__global__ void wait_kernel(const volatile uint32_t* a, uint32_t va, const volatile uint32_t* b, uint32_t vb, uint32_t* err) {
    const long long t0 = clock64();
    bool okA = a == nullptr, okB = b == nullptr;
    while (!(okA && okB)) {
        if (!okA) okA = (int32_t)(*a - va) >= 0;
        if (!okB) okB = (int32_t)(*b - vb) >= 0;
        if (okA && okB) break;
        __nanosleep(200);
        if (clock64() - t0 > 4000000000LL) { *err = 1u; break; }
    }
    __threadfence_system();
}

// The whole halo exchange of one spatial pass as ONE kernel (peer_exchange in romis_gpu.cu):
//   1. (first pass of a frame) wait until the neighbours have finished reading their halo rows of this buffer,
//   2. copy my boundary rows into the neighbours' halo rows -- plain 128-bit stores to IPC-mapped peer memory over NVLink,
//   3. the last block to finish publishes the pass token to both neighbours and then holds the kernel (and with it the
//      stream, i.e. the spatial pass behind it) until the neighbours' tokens have arrived.
// The neighbours run on OTHER GPUs, so waiting cannot starve the writer; spins are bounded (~2 s) and report a timeout.
__device__ __forceinline__ bool spin_until(const volatile uint32_t* f, uint32_t token, uint32_t* err) {
    if (!f) return true;
    const long long t0 = clock64();
    while ((int32_t)(*f - token) < 0) {
        __nanosleep(100);
        if (clock64() - t0 > 4000000000LL) { *err = 1u; return false; }
    }
    return true;
}
__global__ void __launch_bounds__(256) halo_push_kernel(const uint4* __restrict__ src_low, uint4* __restrict__ dst_low, size_t n_low,
                                                        const uint4* __restrict__ src_high, uint4* __restrict__ dst_high, size_t n_high,
                                                        const volatile uint32_t* war_a, const volatile uint32_t* war_b, uint32_t war_token,
                                                        volatile uint32_t* sig_a, volatile uint32_t* sig_b, uint32_t token,
                                                        const volatile uint32_t* wait_a, const volatile uint32_t* wait_b,
                                                        uint32_t* err, unsigned int* ticket) {
    if (threadIdx.x == 0) { spin_until(war_a, war_token, err); spin_until(war_b, war_token, err); }
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = t; i < n_low; i += stride) dst_low[i] = src_low[i];
    for (size_t i = t; i < n_high; i += stride) dst_high[i] = src_high[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int mine = atomicAdd(ticket, 1u);
        if (mine == gridDim.x - 1) {
            __threadfence_system();
            *ticket = 0u;
            if (sig_a) *sig_a = token;
            if (sig_b) *sig_b = token;
            __threadfence_system();
            spin_until(wait_a, token, err);
            spin_until(wait_b, token, err);
            __threadfence_system();
        }
    }
}

// unpack a reservoir buffer into the flat arrays of romis_reservoir_dump (parity read-back)
__global__ void dump_kernel(SceneDev sc, FrameDev fr, ResBuf in, int N, const uint32_t* __restrict__ arch_orig, uint32_t* light, float* u, float* v,
                            float* W, uint32_t* M, float* pos, float* col) {
    int x, y; thread_pixel<false>(x, y);
    y += fr.y0;
    if (x >= fr.W || y >= fr.y1) return;
    for (int j = 0; j < N; j++) {
        uint4 rec = res_rec(in, y - fr.ey0, j)[x];
        size_t i = ((size_t)j * fr.H + y) * fr.W + x;
        // an archived record reports the light it was drawn from (its position / colour below are the light's OLD ones)
        if (light) light[i] = (rec.x != ROMIS_NO_LIGHT && (rec.x & ROMIS_LIGHT_ARCHIVED) && arch_orig) ? arch_orig[rec.x & ~ROMIS_LIGHT_ARCHIVED] : rec.x;
        if (u) u[i] = __uint_as_float(rec.y);
        if (v) v[i] = __uint_as_float(rec.z);
        if (W) W[i] = __uint_as_float(rec.w);
        if (M) M[i] = res_m(in, y - fr.ey0, j)[x];
        if (pos || col) {
            v3 p, c; light_sample(sc, rec.x, __uint_as_float(rec.y), __uint_as_float(rec.z), p, c);
            if (pos) { pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z; }
            if (col) { col[3 * i] = c.x; col[3 * i + 1] = c.y; col[3 * i + 2] = c.z; }
        }
    }
}


void launch_primary(cudaStream_t s, dim3 grid, dim3 block, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, int row0, int row1) {
    launch_pdl(primary_kernel, grid, block, s, sc, fr, g, row0, row1);
}
void launch_shade(cudaStream_t s, dim3 grid, dim3 block, int N, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, const ResBuf& in, float* rgb,
                  const FineDev& fd) {
    ROMIS_DISPATCH_N(N, (launch_pdl(shade_kernel<NT>, grid, block, s, sc, fr, g, in, rgb, fd)));
}
void launch_trace(cudaStream_t s, int n, const SceneDev& sc, const float* o, const float* d, const float* tfar, int any_hit,
                  uint8_t* hit, float* t, float* u, float* v, uint32_t* tri) {
    trace_kernel<<<(n + 127) / 128, 128, 0, s>>>(sc, o, d, tfar, n, any_hit, hit, t, u, v, tri);
}
void launch_row_hits(cudaStream_t s, const GBufDev& g, int W, int H, int n_meshes, uint32_t* rows) {
    row_hits_kernel<<<(H + 7) / 8, 256, 0, s>>>(g, W, H, (uint32_t)n_meshes, rows);
}
void launch_halo_push(cudaStream_t s, const void* src_low, void* dst_low, size_t bytes_low, const void* src_high, void* dst_high, size_t bytes_high,
                      const uint32_t* war_a, const uint32_t* war_b, uint32_t war_token, uint32_t* sig_a, uint32_t* sig_b, uint32_t token,
                      const uint32_t* wait_a, const uint32_t* wait_b, uint32_t* err, unsigned int* ticket) {
    halo_push_kernel<<<32, 256, 0, s>>>((const uint4*)src_low, (uint4*)dst_low, bytes_low / 16, (const uint4*)src_high, (uint4*)dst_high, bytes_high / 16,
                                        war_a, war_b, war_token, sig_a, sig_b, token, wait_a, wait_b, err, ticket);
}
void launch_signal(cudaStream_t s, uint32_t* a, uint32_t va, uint32_t* b, uint32_t vb) {
    if (a || b) signal_kernel<<<1, 1, 0, s>>>(a, va, b, vb);
}
void launch_wait(cudaStream_t s, const uint32_t* a, uint32_t va, const uint32_t* b, uint32_t vb, uint32_t* err) {
    if (a || b) wait_kernel<<<1, 1, 0, s>>>(a, va, b, vb, err);
}
void launch_dump(cudaStream_t s, dim3 grid, dim3 block, const SceneDev& sc, const FrameDev& fr, const ResBuf& in, int N, const uint32_t* arch_orig,
                 uint32_t* light, float* u, float* v, float* W, uint32_t* M, float* pos, float* col) {
    dump_kernel<<<grid, block, 0, s>>>(sc, fr, in, N, arch_orig, light, u, v, W, M, pos, col);
}

// ------------------------------------------------------------------------------------------------
// light edits between frames (romis_upload_lights): keep the history's samples on the lights as they WERE
// ------------------------------------------------------------------------------------------------
// 1. the old records of the edited lights move to their archive slots and the remap table learns where they went
__global__ void light_archive_kernel(const float4* __restrict__ lights, float4* __restrict__ arch, uint32_t* __restrict__ remap,
                                     const uint32_t* __restrict__ idx, const uint32_t* __restrict__ slot, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 6 * n) return;
    const int e = t / 6, q = t - 6 * e;
    arch[6 * (size_t)slot[e] + q] = lights[6 * (size_t)idx[e] + q];
    if (q == 0) remap[idx[e]] = ROMIS_LIGHT_ARCHIVED | slot[e];
}
// 2. every history record that holds an edited light is re-pointed to the archive slot; every archive slot still held by a
//    record is marked (a slot no record holds now can never be held again: only the history survives a frame)
__global__ void history_remap_kernel(ResBuf hist, int lrow0, int rows, int N, const uint32_t* __restrict__ remap, uint32_t n_remap,
                                     uint8_t* __restrict__ mark, uint32_t n_slots) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (x >= hist.W || row >= rows) return;
    for (int j = 0; j < N; j++) {
        uint32_t* rec = reinterpret_cast<uint32_t*>(res_rec(hist, lrow0 + row, j) + x);
        uint32_t li = *rec;
        if (li == ROMIS_NO_LIGHT) continue;
        if (!(li & ROMIS_LIGHT_ARCHIVED)) {
            if (li >= n_remap) continue;
            const uint32_t to = remap[li];
            if (to == ROMIS_NO_LIGHT) continue;
            *rec = li = to;
        }
        const uint32_t s = li & ~ROMIS_LIGHT_ARCHIVED;
        if (s < n_slots) mark[s] = 1;
    }
}
// 3. the remap table goes back to "nothing moved"
__global__ void remap_clear_kernel(uint32_t* __restrict__ remap, const uint32_t* __restrict__ idx, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) remap[idx[t]] = ROMIS_NO_LIGHT;
}
void launch_light_archive(cudaStream_t s, const float4* lights, float4* arch, uint32_t* remap, const uint32_t* idx, const uint32_t* slot, int n,
                          const ResBuf& hist, int lrow0, int rows, int N, uint32_t n_remap, uint8_t* mark, uint32_t n_slots) {
    if (n > 0) light_archive_kernel<<<(6 * n + 255) / 256, 256, 0, s>>>(lights, arch, remap, idx, slot, n);
    if (rows > 0) history_remap_kernel<<<dim3((hist.W + 255) / 256, rows), 256, 0, s>>>(hist, lrow0, rows, N, remap, n_remap, mark, n_slots);
    if (n > 0) remap_clear_kernel<<<(n + 255) / 256, 256, 0, s>>>(remap, idx, n);
}
}  // namespace romis

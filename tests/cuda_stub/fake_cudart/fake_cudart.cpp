// A host-only libcudart.so.12 for tests/test_abi_on_fake_runtime.py: the 32 runtime entry points libpmf references
// (nm -D of a `-cudart shared` build), implemented on host memory.  "Device" memory is malloc'ed host memory, copies and
// memsets are real, streams and events are tokens, kernels are NOT executed -- every launch is logged by kernel name
// (taken from the registration nvcc's host stubs perform) with its grid, block and shared-memory size -- except
// `fill_kernel`, which is carried out so that accumulators have their initial value.  cuTensorMapEncodeTiled (reached
// through cudaGetDriverEntryPoint) records and sanity-checks the descriptor parameters.  With this the whole HOST side of
// the library -- handle state, marshalling, round trips, launch sequences, error paths, allocation balance -- runs on a
// machine without a GPU.  Signatures are checked against the toolkit's own cuda_runtime_api.h at compile time.
#include <cuda_runtime_api.h>
#include <cuda.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {
struct Launch { std::string name; unsigned gx, gy, gz, bx, by, bz; size_t smem; };
struct MapRec { int dtype, rank; uint64_t dim[5], stride[5]; uint32_t box[5]; int swizzle, interleave, oob; uint64_t addr; int rc; };
struct Cfg { dim3 g, b; size_t s; void* st; };
std::mutex mu;
std::map<const void*, std::string> g_kernels;
std::map<void*, size_t> g_live;
std::vector<Launch> g_launches;
std::vector<MapRec> g_maps;
std::vector<Cfg> g_cfg;
struct Fill { std::string match; void* ptr; long n; double value; bool f64; };
std::vector<Fill> g_fills;          // "what a kernel would have produced": buffers filled when a kernel of that name launches
long g_mallocs = 0, g_frees = 0, g_host_allocs = 0, g_host_frees = 0, g_streams = 0, g_events = 0, g_syncs = 0, g_bad_frees = 0;
int g_device = 0, g_device_count = 1, g_cc_major = 10, g_fail_malloc_after = -1;
cudaError_t g_last = cudaSuccess;
char g_fatbin_token;

bool inside_live(const void* p, size_t n) {     // is [p, p + n) inside one live "device" block?
    for (auto& kv : g_live) {
        const char* b = static_cast<const char*>(kv.first);
        if (static_cast<const char*>(p) >= b && static_cast<const char*>(p) + n <= b + kv.second) return true;
    }
    return false;
}
long g_oob_copies = 0;
void check_dev(const void* p, size_t n) { if (n && !inside_live(p, n)) ++g_oob_copies; }
}  // namespace

extern "C" {
// ---- what nvcc's host stubs call ----------------------------------------------------------------------------------
void** __cudaRegisterFatBinary(void*) { return reinterpret_cast<void**>(&g_fatbin_token); }
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
char __cudaInitModule(void**) { return 0; }
void __cudaRegisterFunction(void**, const char* hostFun, char*, const char* deviceName, int, uint3*, uint3*, dim3*, dim3*, int*) {
    std::lock_guard<std::mutex> l(mu);
    g_kernels[hostFun] = deviceName ? deviceName : "?";
}
unsigned __cudaPushCallConfiguration(dim3 g, dim3 b, size_t s, struct CUstream_st* st) {
    std::lock_guard<std::mutex> l(mu);
    g_cfg.push_back({g, b, s, st});
    return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3* g, dim3* b, size_t* s, void* st) {
    std::lock_guard<std::mutex> l(mu);
    if (g_cfg.empty()) return cudaErrorInvalidConfiguration;
    Cfg c = g_cfg.back(); g_cfg.pop_back();
    *g = c.g; *b = c.b; *s = c.s; *static_cast<void**>(st) = c.st;
    return cudaSuccess;
}
cudaError_t cudaLaunchKernel(const void* func, dim3 g, dim3 b, void** args, size_t smem, cudaStream_t) {
    std::lock_guard<std::mutex> l(mu);
    auto it = g_kernels.find(func);
    std::string name = it == g_kernels.end() ? "<unregistered>" : it->second;
    g_launches.push_back({name, g.x, g.y, g.z, b.x, b.y, b.z, smem});
    for (const Fill& f : g_fills)
        if (name.find(f.match) != std::string::npos) {
            if (f.f64) for (long i = 0; i < f.n; ++i) static_cast<double*>(f.ptr)[i] = f.value;
            else for (long i = 0; i < f.n; ++i) static_cast<float*>(f.ptr)[i] = (float)f.value;
        }
    if (name.find("fill_kernel") != std::string::npos) {             // fill_kernel(float* p, size_t n, float v)
        float* p = *static_cast<float**>(args[0]);
        size_t n = *static_cast<size_t*>(args[1]);
        float v = *static_cast<float*>(args[2]);
        check_dev(p, n * 4);
        for (size_t i = 0; i < n; ++i) p[i] = v;
    }
    if (g.x == 0 || g.y == 0 || g.z == 0 || b.x * b.y * b.z == 0 || b.x * b.y * b.z > 1024 || smem > 232448) return g_last = cudaErrorInvalidConfiguration;
    return cudaSuccess;
}
// ---- devices ----------------------------------------------------------------------------------------------------------
cudaError_t cudaGetDeviceCount(int* n) { *n = g_device_count; return g_device_count > 0 ? cudaSuccess : cudaErrorNoDevice; }
cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= g_device_count) return g_last = cudaErrorInvalidDevice; g_device = d; return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = g_device; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, enum cudaDeviceAttr a, int) {
    switch (a) {
    case cudaDevAttrMultiProcessorCount: *v = 148; break;
    case cudaDevAttrComputeCapabilityMajor: *v = g_cc_major; break;
    case cudaDevAttrComputeCapabilityMinor: *v = 0; break;
    case cudaDevAttrMaxSharedMemoryPerBlockOptin: *v = 232448; break;
    default: *v = 0;
    }
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize(void) { ++g_syncs; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { cudaError_t e = g_last; g_last = cudaSuccess; return e; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "fake runtime error"; }
cudaError_t cudaFuncSetAttribute(const void*, enum cudaFuncAttribute, int) { return cudaSuccess; }
// ---- memory -----------------------------------------------------------------------------------------------------------
cudaError_t cudaMalloc(void** p, size_t n) {
    std::lock_guard<std::mutex> l(mu);
    if (g_fail_malloc_after == 0) { *p = nullptr; return g_last = cudaErrorMemoryAllocation; }
    if (g_fail_malloc_after > 0) --g_fail_malloc_after;
    void* q = std::malloc(n ? n : 1);
    if (!q) { *p = nullptr; return g_last = cudaErrorMemoryAllocation; }
    std::memset(q, 0xCD, n);                                          // fresh device memory is not zero
    g_live[q] = n; ++g_mallocs; *p = q;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    std::lock_guard<std::mutex> l(mu);
    if (!p) return cudaSuccess;
    auto it = g_live.find(p);
    if (it == g_live.end()) { ++g_bad_frees; return g_last = cudaErrorInvalidValue; }
    g_live.erase(it); ++g_frees; std::free(p);
    return cudaSuccess;
}
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = std::calloc(1, n ? n : 1); ++g_host_allocs; return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { if (p) { std::free(p); ++g_host_frees; } return cudaSuccess; }
static cudaError_t do_copy(void* d, const void* s, size_t n, cudaMemcpyKind k) {
    std::lock_guard<std::mutex> l(mu);
    if (k == cudaMemcpyHostToDevice || k == cudaMemcpyDeviceToDevice) check_dev(d, n);
    if (k == cudaMemcpyDeviceToHost || k == cudaMemcpyDeviceToDevice) check_dev(s, n);
    std::memmove(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind k) { return do_copy(d, s, n, k); }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t) { return do_copy(d, s, n, k); }
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t) {
    if (w > dp || w > sp) return g_last = cudaErrorInvalidPitchValue;
    for (size_t r = 0; r < h; ++r) do_copy(static_cast<char*>(d) + r * dp, static_cast<const char*>(s) + r * sp, w, k);
    return cudaSuccess;
}
cudaError_t cudaMemset(void* p, int v, size_t n) { std::lock_guard<std::mutex> l(mu); check_dev(p, n); std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { return cudaMemset(p, v, n); }
// ---- streams, events ------------------------------------------------------------------------------------------------------
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(std::malloc(8)); ++g_streams; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); --g_streams; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { ++g_syncs; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = reinterpret_cast<cudaEvent_t>(std::malloc(8)); ++g_events; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); --g_events; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

// ---- the one driver entry point the library asks for ---------------------------------------------------------------------
static CUresult fake_encode_tiled(CUtensorMap* m, CUtensorMapDataType dt, cuuint32_t rank, void* addr, const cuuint64_t* dim,
                                  const cuuint64_t* stride, const cuuint32_t* box, const cuuint32_t* estride,
                                  CUtensorMapInterleave il, CUtensorMapSwizzle sw, CUtensorMapL2promotion, CUtensorMapFloatOOBfill oob) {
    std::lock_guard<std::mutex> l(mu);
    MapRec r{};
    r.dtype = dt; r.rank = (int)rank; r.swizzle = sw; r.interleave = il; r.oob = oob; r.addr = (uint64_t)addr; r.rc = 0;
    const int esz = (dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 || dt == CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 || dt == CU_TENSOR_MAP_DATA_TYPE_INT32 ||
                     dt == CU_TENSOR_MAP_DATA_TYPE_UINT32) ? 4 : (dt == CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 || dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT16 ||
                     dt == CU_TENSOR_MAP_DATA_TYPE_UINT16) ? 2 : (dt == CU_TENSOR_MAP_DATA_TYPE_UINT8 ? 1 : 8);
    // the constraints the driver documents for cuTensorMapEncodeTiled
    if (rank < 1 || rank > 5 || ((uint64_t)addr & 15) || !m) r.rc = 1;
    for (cuuint32_t i = 0; i < rank && i < 5; ++i) {
        r.dim[i] = dim[i]; r.box[i] = box[i];
        if (dim[i] == 0 || dim[i] > (1ull << 32) || box[i] == 0 || box[i] > 256 || estride[i] == 0 || estride[i] > 8) r.rc = 1;
        if (i + 1 < rank) { r.stride[i] = stride[i]; if ((stride[i] & 15) || stride[i] >= (1ull << 40)) r.rc = 1; }
    }
    const uint64_t inner = (uint64_t)box[0] * esz;
    if (inner % 16) r.rc = 1;
    if (sw == CU_TENSOR_MAP_SWIZZLE_32B && inner > 32) r.rc = 1;
    if (sw == CU_TENSOR_MAP_SWIZZLE_64B && inner > 64) r.rc = 1;
    if (sw >= CU_TENSOR_MAP_SWIZZLE_128B && inner > 128) r.rc = 1;
    if (rank >= 2 && !inside_live(addr, 1)) r.rc = 2;                 // the tensor must be device memory
    if (r.rc == 0) {                                                  // ... and so must all of it: TMA clamps to the descriptor's dims,
        uint64_t extent = dim[0] * (uint64_t)esz;                     // not to the allocation, so a descriptor larger than its block
        for (cuuint32_t i = 1; i < rank; ++i) extent += (dim[i] - 1) * stride[i - 1];   // lets loads / reduce-adds run past it
        if (!inside_live(addr, extent)) r.rc = 3;
    }
    g_maps.push_back(r);
    if (m) std::memset(m, 0, sizeof(CUtensorMap));
    return r.rc ? CUDA_ERROR_INVALID_VALUE : CUDA_SUCCESS;
}
cudaError_t cudaGetDriverEntryPoint(const char* sym, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* st) {
    if (std::strcmp(sym, "cuTensorMapEncodeTiled") == 0) {
        *fn = reinterpret_cast<void*>(&fake_encode_tiled);
        if (st) *st = cudaDriverEntryPointSuccess;
        return cudaSuccess;
    }
    *fn = nullptr;
    if (st) *st = cudaDriverEntryPointSymbolNotFound;
    return cudaErrorInvalidValue;
}

// ---- doors for the test --------------------------------------------------------------------------------------------------
void fake_counters(long* out) {       // mallocs, frees, live blocks, live bytes, host allocs, host frees, streams, events, bad frees, oob copies
    std::lock_guard<std::mutex> l(mu);
    size_t bytes = 0; for (auto& kv : g_live) bytes += kv.second;
    long v[10] = {g_mallocs, g_frees, (long)g_live.size(), (long)bytes, g_host_allocs, g_host_frees, g_streams, g_events, g_bad_frees, g_oob_copies};
    std::memcpy(out, v, sizeof v);
}
long fake_syncs(void) { std::lock_guard<std::mutex> l(mu); return g_syncs; }      // cudaStreamSynchronize + cudaDeviceSynchronize calls so far
int fake_launch_count(void) { std::lock_guard<std::mutex> l(mu); return (int)g_launches.size(); }
int fake_launch(int i, char* name, int cap, unsigned* dims, size_t* smem) {
    std::lock_guard<std::mutex> l(mu);
    if (i < 0 || i >= (int)g_launches.size()) return -1;
    const Launch& x = g_launches[i];
    std::strncpy(name, x.name.c_str(), cap - 1); name[cap - 1] = 0;
    unsigned d[6] = {x.gx, x.gy, x.gz, x.bx, x.by, x.bz}; std::memcpy(dims, d, sizeof d); *smem = x.smem;
    return 0;
}
void fake_clear_launches(void) { std::lock_guard<std::mutex> l(mu); g_launches.clear(); g_maps.clear(); }
int fake_map_count(void) { std::lock_guard<std::mutex> l(mu); return (int)g_maps.size(); }
int fake_map(int i, long long* out) {   // dtype, rank, swizzle, oob, rc, dim0, dim1, box0, box1, stride0
    std::lock_guard<std::mutex> l(mu);
    if (i < 0 || i >= (int)g_maps.size()) return -1;
    const MapRec& r = g_maps[i];
    long long v[10] = {r.dtype, r.rank, r.swizzle, r.oob, r.rc, (long long)r.dim[0], (long long)r.dim[1], r.box[0], r.box[1], (long long)r.stride[0]};
    std::memcpy(out, v, sizeof v);
    return 0;
}
void fake_fill_on_launch(const char* match, void* ptr, long n, double value, int f64) {
    std::lock_guard<std::mutex> l(mu);
    g_fills.push_back({match, ptr, n, value, f64 != 0});
}
void fake_clear_fills(void) { std::lock_guard<std::mutex> l(mu); g_fills.clear(); }
void fake_set(int device_count, int cc_major, int fail_malloc_after) { g_device_count = device_count; g_cc_major = cc_major; g_fail_malloc_after = fail_malloc_after; }
}

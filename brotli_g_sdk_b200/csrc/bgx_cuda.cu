// bgx_cuda.cu -- CUDA kernels + C ABI (include/brotlig_b200.h) of the B200-native Brotli-G decompressor.
//
// Replaces, for the decode path only:
//   src/decoder/BrotliGCompute.hlsl (CSMain :1753-1881: persistent waves pulling pages with atomics)
//       -> bgx_decode_pages_kernel: persistent two-warp CTAs (producer + consumer per page), one atomic page counter over ALL streams
//   sample/BrotligGPUDecoder.cpp (DecodeGPU :260-748: upload, Dispatch, readback, timestamp queries)
//       -> bgx_decode_host / bgx_decode_batch_host / bgx_plan_*: CUDA stream + events
// Build: nvcc -gencode arch=compute_100a,code=sm_100a (see brotli_g_sdk_b200/build.py).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/brotlig_b200.h"
#include "bgx_format.h"
#include "host_plan.h"
#include "page_decode.cuh"

namespace {

using bgx::PreconLayout;
using bgx::StreamInfo;

// ------------------------------------------------------------------------------------ device side
struct StreamDev {
  const uint8_t* table;     // page table (device)
  const uint8_t* pages;     // first page byte
  const uint8_t* src_end;   // one past the last readable byte of the stream buffer
  uint8_t* dst;             // where page `page_begin` goes (texture streams: the conditioned scratch planes)
  uint32_t num_pages, page_size, last_page_size;
  uint32_t page_begin;      // first page of the stream this plan decodes
  uint32_t first_q;         // position of that page in the flat work queue
  uint32_t allow_delta;
  bgxk::DeltaPlanes planes; // colour planes for the per-page delta decode
};

struct QueueCtl {
  uint32_t next_page;       // atomic work counter (cf. meta.InterlockedAdd, BrotliGCompute.hlsl:1815)
  uint32_t bad_pages;       // pages whose decode reported an error
};

constexpr int kPageThreads = 64;   // two warps per page: producer (entropy decode) + consumer (LZ77 assembly)

#ifndef BGX_CTAS_PER_SM
#define BGX_CTAS_PER_SM 16
#endif
__global__ void __launch_bounds__(kPageThreads, BGX_CTAS_PER_SM) bgx_decode_pages_kernel(const StreamDev* __restrict__ streams, uint32_t nstreams,
                                                                        uint32_t q_begin, uint32_t q_end, float q_to_stream,
                                                                        QueueCtl* ctl, uint32_t* __restrict__ page_status) {
  __shared__ bgxk::WarpSmem sm;
  uint32_t& q_shared = sm.q_shared;
  const bool layout_ok = bgxk::arena_layout_ok(sm.ring, sm.litq);
  const uint32_t tid = threadIdx.x;
  uint32_t cur_stream = 0;   // owner of the last page this CTA decoded (streams are in queue order)
  bool first_compressed = true;   // the page arena's mbarriers have not been initialised yet
  if (tid == 0) q_shared = atomicAdd(&ctl->next_page, 1u);
  for (;;) {
    __syncthreads();
    const uint32_t q = q_begin + q_shared;   // [q_begin, q_end): the slice of the flat page queue this launch owns
    if (q >= q_end) break;
    uint32_t q_next = 0;
    if (tid == 0) q_next = atomicAdd(&ctl->next_page, 1u);   // claim the next page now: the atomic's latency hides behind this page
    // stream owning queue slot q: last stream with first_q <= q. A CTA claims increasing slots, so the owner is
    // almost always the previous page's stream or its successor: one load decides, the search only runs beyond.
    uint32_t lo = cur_stream;
    if (lo + 1 < nstreams && streams[lo + 1].first_q <= q) {
      ++lo;
      uint32_t hi = nstreams;
      // batches of many small streams: a proportional guess (exact when the streams hold equal numbers of pages)
      // brackets the owner with two independent loads, so the search below rarely runs
      uint32_t g = (uint32_t)((float)q * q_to_stream);
      g = g < nstreams - 1u ? g : nstreams - 1u;
      const uint32_t fg = streams[g].first_q, fg1 = g + 1 < nstreams ? streams[g + 1].first_q : 0xffffffffu;
      if (fg <= q) {
        lo = g > lo ? g : lo;
        if (fg1 > q) hi = g + 1u;
      } else {
        hi = g;
      }
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (streams[mid].first_q <= q) lo = mid; else hi = mid;
      }
    }
    cur_stream = lo;
    const StreamDev& s = streams[lo];
    const uint32_t page = s.page_begin + (q - s.first_q);
    StreamInfo si;
    si.num_pages = s.num_pages;
    si.page_size = s.page_size;
    si.last_page_size = s.last_page_size;
    // page table entries as aligned 32-bit loads (the stream is 16-byte aligned, its headers are 8 or 16 bytes);
    // same arithmetic as bgx::page_extent (host_plan.h)
    bgx::PageExtent e;
    {
      const uint32_t* tbl = reinterpret_cast<const uint32_t*>(s.table);
      const uint32_t t_next = tbl[page + 1 < si.num_pages ? page + 1 : 0];   // entry 0 holds the size of the last page
      e.in_off = page ? tbl[page] : 0u;
      e.in_size = (page + 1 < si.num_pages) ? t_next - e.in_off : t_next;
      e.out_off = page * si.page_size;
      e.out_size = (page + 1 == si.num_pages && si.last_page_size) ? si.last_page_size : si.page_size;
    }
    const uint8_t* in = s.pages + e.in_off;
    uint8_t* out = s.dst + (size_t)(page - s.page_begin) * s.page_size;
    uint32_t status = 0;
    // a corrupt page table must not send any access outside the stream buffer (raw pages included), and the
    // bit readers need their page on a 4-byte boundary (the encoder pads pages to 4 bytes)
    const size_t avail = (size_t)(s.src_end - s.pages);
    const bool in_bounds = (size_t)e.in_off <= avail && (size_t)e.in_size <= avail - (size_t)e.in_off;
    if (!in_bounds) {
      status = bgxk::kPageErrTable;
    } else if (e.in_size == e.out_size) {
      bgxk::copy_page_cta(out, in, e.out_size, &sm);
    } else if (e.in_size < 8u || (e.in_off & 3u) != 0u) {
      status = bgxk::kPageErrTable;
    } else if (!layout_ok) {
      status = bgxk::kPageErrLayout;
    } else {
      bgxk::PageJob job;
      job.in = in;
      job.in_size = e.in_size;
      const size_t room = avail - (size_t)e.in_off;
      job.in_limit = room > 0xfffffff0u ? 0xfffffff0u : (uint32_t)room;
      job.out = out;
      job.out_size = e.out_size;
      job.allow_delta = s.allow_delta;
      const bgxk::PageResult r = bgxk::decode_page_cta(job, &sm, first_compressed);
      first_compressed = false;
      status = r.status;
      if (!status && r.is_delta) {
        if (tid >= 32) bgxk::delta_decode_warp(out, e.out_off, e.out_size, s.planes);   // the consumer warp wrote the page
        status |= 0x40000000u;   // informational: page was delta coded
      }
    }
    if (tid == 0) {
      page_status[q] = status;
      if (status & 0xffffu) atomicAdd(&ctl->bad_pages, 1u);
    }
    __syncthreads();   // everybody is done with q_shared and the page arena
    if (tid == 0) q_shared = q_next;
  }
}

struct PreconDev {          // one pre-conditioned stream of a launch
  const PreconLayout* layout;
  const uint8_t* planes;    // conditioned scratch planes (what the page kernel wrote)
  uint8_t* tex;             // the caller's output
};

// one thread per texture block (blockIdx.y = stream): gather the block's fields from the planes, write the block
template <int FMT>
__device__ __forceinline__ void decondition_stream_fast(const PreconLayout& L, const uint8_t* planes, uint8_t* tex) {
  const uint32_t total = L.total_blocks;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
    bgxk::decondition_block_fast<FMT>(L, t, planes, tex);
}

__global__ void __launch_bounds__(256) bgx_decondition_kernel(const PreconDev* __restrict__ jobs) {
  const PreconDev j = jobs[blockIdx.y];
  const PreconLayout& L = *j.layout;
  // vector path: planes on a 8-byte boundary, every block of every mip on a block_bytes boundary
  bool aligned = ((reinterpret_cast<uintptr_t>(j.planes) & 7u) == 0) &&
                 ((reinterpret_cast<uintptr_t>(j.tex) & (L.block_bytes - 1u)) == 0);
  for (uint32_t m = 0; m < L.num_mips; ++m)
    aligned = aligned && ((L.mip_off_bytes[m] | L.pitch_bytes[m]) & (L.block_bytes - 1u)) == 0;
  // (plane starts are multiples of total_blocks; 8-byte fields additionally need an 8-aligned plane start, which
  //  holds for BC2's first plane at offset 0)
  if (aligned) {
    switch (L.format) {
      case 1: decondition_stream_fast<1>(L, j.planes, j.tex); return;
      case 2: decondition_stream_fast<2>(L, j.planes, j.tex); return;
      case 3: decondition_stream_fast<3>(L, j.planes, j.tex); return;
      case 4: decondition_stream_fast<4>(L, j.planes, j.tex); return;
      case 5: decondition_stream_fast<5>(L, j.planes, j.tex); return;
      default: break;
    }
  }
  const uint32_t total = L.total_blocks;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
    bgxk::decondition_block(L, t, j.planes, j.tex);
}

}  // namespace

// ------------------------------------------------------------------------------------ host side
struct bgx_context {
  int device = 0;
  cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;   // kernels / uploads / downloads
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = 0;
  int blocks_per_sm = 0;
  std::string err;
  // grow-only device arenas for the host-pointer entry points
  uint8_t* d_in = nullptr;
  size_t d_in_cap = 0;
  uint8_t* d_out = nullptr;
  size_t d_out_cap = 0;
  uint8_t* d_scratch = nullptr;   // conditioned planes of texture streams decoded through the host-pointer calls
  size_t d_scratch_cap = 0;
  uint8_t* d_meta = nullptr;      // work descriptors / queue control / page status of the host-pointer calls' plans
  size_t d_meta_cap = 0;
  std::vector<cudaEvent_t> ev_pool;   // timing events of the host-pointer pipeline (created once, reused by every call)
  cudaEvent_t ev_start = nullptr;
};

struct PreconJob {
  uint32_t stream_index = 0;      // position of the stream in the caller's list
  PreconLayout layout;
  PreconLayout* d_layout = nullptr;
  uint8_t* d_planes = nullptr;    // conditioned scratch (decode target)
  uint8_t* d_tex = nullptr;       // final output
  uint32_t out_size = 0;
  bool has_padding = false;
};

struct bgx_plan {
  std::vector<StreamDev> h_streams;
  StreamDev* d_streams = nullptr;
  QueueCtl* d_ctl = nullptr;
  uint32_t* d_status = nullptr;
  uint32_t total_pages = 0;
  std::vector<PreconJob> precon;
  uint8_t* d_scratch = nullptr;   // backing store of all conditioned scratch planes
  bool owns_scratch = true;       // false: borrowed from the context's grow-only arena (host-pointer calls)
  bool owns_meta = true;          // false: d_streams / d_ctl / d_status / d_layouts / d_precon are carved from ctx->d_meta
  PreconDev* d_precon = nullptr;  // device copy of the pre-conditioned jobs, in stream order
  PreconLayout* d_layouts = nullptr;   // their layouts, one array
  bgx_plan_info info{};
  cudaStream_t last_stream = nullptr;
  std::vector<uint32_t> q_start;  // [n+1]: first queue slot of caller stream i (prefix sum of its page count)
  uint32_t groups_used = 1;
};
constexpr uint32_t kMaxGroups = 64;

namespace {

int fail(bgx_context* ctx, const char* what, cudaError_t e) {
  if (ctx) ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
  return bgx::kErrGeneric;
}
#define BGX_CUDA(ctx, call)                                  \
  do {                                                       \
    cudaError_t e_ = (call);                                 \
    if (e_ != cudaSuccess) return fail(ctx, #call, e_);      \
  } while (0)

int grow(bgx_context* ctx, uint8_t** p, size_t* cap, size_t need) {
  if (need <= *cap) return 0;
  if (*p) BGX_CUDA(ctx, cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  const size_t want = need + need / 8 + 4096;
  BGX_CUDA(ctx, cudaMalloc(p, want));
  *cap = want;
  return 0;
}

}  // namespace

extern "C" {

uint32_t bgx_decompressed_size(const uint8_t* src) {
  StreamInfo si;
  // DecompressedSize does not validate (BrotligDecoder.cpp:34-38); neither do we.
  const uint32_t w0 = bgx::load_le32(src), w1 = bgx::load_le32(src + 4);
  si.num_pages = w0 >> 16;
  si.page_size = bgx::kMinPageSize << (w1 & 3u);
  si.last_page_size = (w1 >> 2) & 0x3ffffu;
  return si.num_pages * si.page_size - (si.last_page_size ? si.page_size - si.last_page_size : 0u);
}

int bgx_create(bgx_context** out, int device) {
  *out = nullptr;
  bgx_context* ctx = new bgx_context();
  cudaError_t e;
  if (device < 0) {
    e = cudaGetDevice(&device);
    if (e != cudaSuccess) { fprintf(stderr, "brotlig_b200: no CUDA device: %s\n", cudaGetErrorString(e)); delete ctx; return bgx::kErrGeneric; }
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) { fprintf(stderr, "brotlig_b200: cudaSetDevice(%d): %s\n", device, cudaGetErrorString(e)); delete ctx; return bgx::kErrGeneric; }
  ctx->device = device;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { fprintf(stderr, "brotlig_b200: cudaGetDeviceProperties: %s\n", cudaGetErrorString(e)); delete ctx; return bgx::kErrGeneric; }
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
    fprintf(stderr, "brotlig_b200: stream/event creation failed\n");
    delete ctx;
    return bgx::kErrGeneric;
  }
  // two-warp CTAs, as many as registers and the shared-memory arena allow per SM (16)
  cudaFuncSetAttribute(bgx_decode_pages_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bgx_decode_pages_kernel, kPageThreads, 0);
  if (e != cudaSuccess || per_sm < 1) {
    fprintf(stderr, "brotlig_b200: kernel image not usable on this device (%s); built for sm_100a\n", cudaGetErrorString(e));
    delete ctx;
    return bgx::kErrGeneric;
  }
  ctx->blocks_per_sm = per_sm;
  if (const char* cap = getenv("BGX_CTAS_CAP")) {   // kernel experiments: fewer resident pages per SM
    const int c = atoi(cap);
    if (c >= 1 && c < per_sm) ctx->blocks_per_sm = c;
  }
  *out = ctx;
  return bgx::kOk;
}

void bgx_destroy(bgx_context* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->d_in) cudaFree(ctx->d_in);
  if (ctx->d_out) cudaFree(ctx->d_out);
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->d_meta) cudaFree(ctx->d_meta);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
  if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
  delete ctx;
}

const char* bgx_last_error(const bgx_context* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

void bgx_plan_destroy(bgx_plan* plan) {
  if (!plan) return;
  if (plan->owns_meta) {
    if (plan->d_streams) cudaFree(plan->d_streams);
    if (plan->d_ctl) cudaFree(plan->d_ctl);
    if (plan->d_status) cudaFree(plan->d_status);
    if (plan->d_precon) cudaFree(plan->d_precon);
    if (plan->d_layouts) cudaFree(plan->d_layouts);
  }
  if (plan->d_scratch && plan->owns_scratch) cudaFree(plan->d_scratch);
  delete plan;
}

void bgx_plan_get_info(const bgx_plan* plan, bgx_plan_info* info) { *info = plan->info; }

static int plan_create_impl(bgx_context* ctx, const bgx_stream* streams, uint32_t n, bgx_plan** out, bool ctx_scratch) {
  *out = nullptr;
  BGX_CUDA(ctx, cudaSetDevice(ctx->device));
  bgx_plan* plan = new bgx_plan();
  struct Guard { bgx_plan* p; ~Guard() { if (p) bgx_plan_destroy(p); } } guard{plan};
  uint64_t scratch_bytes = 0;
  std::vector<size_t> scratch_off;
  plan->q_start.assign(n + 1, 0);
  for (uint32_t i = 0; i < n; ++i) {
    plan->q_start[i] = plan->total_pages;
    plan->q_start[i + 1] = plan->total_pages;
    const bgx_stream& s = streams[i];
    StreamInfo si;
    const int rc = bgx::parse_stream_header(s.header, &si);
    if (rc) { ctx->err = "stream " + std::to_string(i) + ": bad header"; return rc; }
    if ((reinterpret_cast<uintptr_t>(s.d_src) & 15u) != 0) { ctx->err = "stream pointer must be 16-byte aligned"; return bgx::kErrGeneric; }
    const uint64_t table_end = (uint64_t)si.header_bytes + 4ull * si.num_pages;
    if (s.src_capacity < ((s.src_size + 15u) & ~15u)) {
      ctx->err = "src_capacity must cover the stream rounded up to 16 bytes (the input is staged in 16-byte chunks)";
      return bgx::kErrGeneric;
    }
    if (table_end > s.src_size || s.src_capacity < s.src_size) { ctx->err = "stream " + std::to_string(i) + ": truncated"; return bgx::kErrCorruptStream; }
    uint32_t begin = s.page_begin, count = s.page_count ? s.page_count : (si.num_pages > begin ? si.num_pages - begin : 0);
    if (begin > si.num_pages || count > si.num_pages - begin) { ctx->err = "page range outside the stream"; return bgx::kErrGeneric; }
    if (count == 0) continue;
    const bool whole = begin == 0 && count == si.num_pages;
    // bytes this range produces
    uint64_t produced = (uint64_t)count * si.page_size;
    if (begin + count == si.num_pages && si.last_page_size) produced -= si.page_size - si.last_page_size;
    if (produced > s.dst_capacity) { ctx->err = "stream " + std::to_string(i) + ": output buffer too small"; return bgx::kErrGeneric; }

    StreamDev d{};
    d.table = s.d_src + si.header_bytes;
    d.pages = d.table + 4ull * si.num_pages;
    d.src_end = s.d_src + s.src_capacity;
    d.dst = s.d_dst;
    d.num_pages = si.num_pages;
    d.page_size = si.page_size;
    d.last_page_size = si.last_page_size;
    d.page_begin = begin;
    d.first_q = plan->total_pages;
    d.allow_delta = si.preconditioned;
    if (si.preconditioned) {
      if (!whole) { ctx->err = "page ranges of pre-conditioned streams are not supported"; return bgx::kErrGeneric; }
      const bgx::PreconHeaderFields f = bgx::parse_precon_header(s.header + 8);
      PreconJob pj;
      // BrotligDecoder.cpp:478 sizes the layout from the caller's buffer size; we require the exact size
      if (!bgx::precon_layout_init(&pj.layout, f.format, f.width_blocks, f.height_blocks, f.pitch_bytes, f.num_mips,
                                   f.swizzled != 0, f.pitch_aligned != 0, si.uncompressed_size)) {
        ctx->err = "stream " + std::to_string(i) + ": texture layout does not match the stream size";
        return bgx::kErrCorruptStream;
      }
      pj.d_tex = s.d_dst;
      pj.out_size = si.uncompressed_size;
      pj.has_padding = (uint64_t)pj.layout.total_blocks * pj.layout.block_bytes != si.uncompressed_size;
      d.planes.count = pj.layout.num_color_sub;
      for (uint32_t c = 0; c < pj.layout.num_color_sub; ++c) {
        d.planes.lo[c] = pj.layout.sub_stream_off[pj.layout.color_sub[c]];
        d.planes.hi[c] = pj.layout.sub_stream_off[pj.layout.color_sub[c] + 1];
      }
      scratch_off.push_back((size_t)scratch_bytes);
      scratch_bytes += ((uint64_t)si.uncompressed_size + 255u) & ~255ull;
      pj.stream_index = i;
      plan->precon.push_back(pj);
      d.dst = nullptr;   // patched below once the scratch arena exists
      d.allow_delta = 1u | ((uint32_t)plan->precon.size() << 8);   // remember which precon job (index+1) in the high bits
    }
    plan->h_streams.push_back(d);
    plan->total_pages += count;
    plan->q_start[i + 1] = plan->total_pages;
    plan->info.compressed_bytes += s.src_size;   // whole-stream bytes; page ranges read a subset (reported as upper bound)
    plan->info.uncompressed_bytes += produced;
  }
  plan->info.pages = plan->total_pages;
  // device memory of the plan: per-plan allocations, or -- for the host-pointer calls, which build a plan per call --
  // one grow-only arena of the context, so that a call costs no cudaMalloc once the arenas are warm
  const size_t ns = std::max<size_t>(plan->h_streams.size(), 1);
  const size_t np = plan->precon.size();
  auto up256 = [](size_t n) { return (n + 255u) & ~(size_t)255u; };
  const size_t sz_streams = up256(ns * sizeof(StreamDev)), sz_ctl = up256(kMaxGroups * sizeof(QueueCtl));
  const size_t sz_status = up256(std::max<size_t>(plan->total_pages, 1) * sizeof(uint32_t));
  const size_t sz_layouts = up256(np * sizeof(PreconLayout)), sz_precon = up256(np * sizeof(PreconDev));
  if (ctx_scratch) {
    if (grow(ctx, &ctx->d_meta, &ctx->d_meta_cap, sz_streams + sz_ctl + sz_status + sz_layouts + sz_precon)) return bgx::kErrGeneric;
    uint8_t* p = ctx->d_meta;
    plan->owns_meta = false;
    plan->d_streams = reinterpret_cast<StreamDev*>(p); p += sz_streams;
    plan->d_ctl = reinterpret_cast<QueueCtl*>(p); p += sz_ctl;
    plan->d_status = reinterpret_cast<uint32_t*>(p); p += sz_status;
    if (np) { plan->d_layouts = reinterpret_cast<PreconLayout*>(p); p += sz_layouts; plan->d_precon = reinterpret_cast<PreconDev*>(p); }
  } else {
    BGX_CUDA(ctx, cudaMalloc(&plan->d_streams, sz_streams));
    BGX_CUDA(ctx, cudaMalloc(&plan->d_ctl, sz_ctl));
    BGX_CUDA(ctx, cudaMalloc(&plan->d_status, sz_status));
    if (np) {
      BGX_CUDA(ctx, cudaMalloc(&plan->d_layouts, sz_layouts));
      BGX_CUDA(ctx, cudaMalloc(&plan->d_precon, sz_precon));
    }
  }
  if (scratch_bytes) {
    if (ctx_scratch) {
      if (grow(ctx, &ctx->d_scratch, &ctx->d_scratch_cap, scratch_bytes)) return bgx::kErrGeneric;
      plan->d_scratch = ctx->d_scratch;
      plan->owns_scratch = false;
    } else {
      BGX_CUDA(ctx, cudaMalloc(&plan->d_scratch, scratch_bytes));
    }
    for (auto& d : plan->h_streams) {
      if (d.allow_delta >> 8) {
        const size_t k = (d.allow_delta >> 8) - 1;
        plan->precon[k].d_planes = plan->d_scratch + scratch_off[k];
        d.dst = plan->precon[k].d_planes;
        d.allow_delta = 1;
      }
    }
    // all texture layouts of the plan in one device array (one copy)
    std::vector<PreconLayout> hl;
    for (auto& p : plan->precon) hl.push_back(p.layout);
    BGX_CUDA(ctx, cudaMemcpy(plan->d_layouts, hl.data(), hl.size() * sizeof(PreconLayout), cudaMemcpyHostToDevice));
    std::vector<PreconDev> hj;
    for (size_t k = 0; k < plan->precon.size(); ++k) {
      PreconJob& p = plan->precon[k];
      p.d_layout = plan->d_layouts + k;
      hj.push_back(PreconDev{p.d_layout, p.d_planes, p.d_tex});
    }
    BGX_CUDA(ctx, cudaMemcpy(plan->d_precon, hj.data(), hj.size() * sizeof(PreconDev), cudaMemcpyHostToDevice));
  }
  if (!plan->h_streams.empty())
    BGX_CUDA(ctx, cudaMemcpy(plan->d_streams, plan->h_streams.data(), plan->h_streams.size() * sizeof(StreamDev), cudaMemcpyHostToDevice));
  plan->info.kernels_per_launch = (plan->total_pages ? 1u : 0u) + (plan->precon.empty() ? 0u : 1u);
  plan->info.sm_count = (uint32_t)ctx->sm_count;
  plan->info.block_threads = kPageThreads;
  plan->info.smem_bytes_per_block = (uint32_t)sizeof(bgxk::WarpSmem);
  const uint64_t max_blocks = (uint64_t)ctx->sm_count * ctx->blocks_per_sm;
  plan->info.grid_blocks = (uint32_t)std::min<uint64_t>(max_blocks, std::max<uint32_t>(plan->total_pages, 1));
  guard.p = nullptr;
  *out = plan;
  return bgx::kOk;
}

int bgx_plan_create(bgx_context* ctx, const bgx_stream* streams, uint32_t n, bgx_plan** out) {
  return plan_create_impl(ctx, streams, n, out, false);
}

// Enqueues the decode of caller streams [a, b) on `st`, using work-queue control block `group`.
static int launch_range(bgx_context* ctx, bgx_plan* plan, uint32_t a, uint32_t b, uint32_t group, cudaStream_t st) {
  const uint32_t q0 = plan->q_start[a], q1 = plan->q_start[b];
  BGX_CUDA(ctx, cudaMemsetAsync(plan->d_ctl + group, 0, sizeof(QueueCtl), st));
  for (auto& p : plan->precon)
    if (p.stream_index >= a && p.stream_index < b && p.has_padding)
      BGX_CUDA(ctx, cudaMemsetAsync(p.d_tex, 0, p.out_size, st));   // pitch padding stays 0 (BrotligDecoder.cpp:448)
  if (q1 > q0) {
    const uint64_t max_blocks = (uint64_t)ctx->sm_count * ctx->blocks_per_sm;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(max_blocks, q1 - q0);
    const float q_to_stream = (float)plan->h_streams.size() / (float)std::max<uint32_t>(plan->total_pages, 1u);
    bgx_decode_pages_kernel<<<grid, kPageThreads, 0, st>>>(plan->d_streams, (uint32_t)plan->h_streams.size(), q0, q1, q_to_stream,
                                                 plan->d_ctl + group, plan->d_status);
    BGX_CUDA(ctx, cudaGetLastError());
  }
  // all pre-conditioned streams of the range in ONE launch (they are contiguous in plan->precon: stream order)
  size_t j0 = plan->precon.size(), j1 = 0;
  uint32_t max_blocks = 0;
  for (size_t k = 0; k < plan->precon.size(); ++k) {
    const PreconJob& p = plan->precon[k];
    if (p.stream_index < a || p.stream_index >= b) continue;
    j0 = std::min(j0, k);
    j1 = std::max(j1, k + 1);
    max_blocks = std::max(max_blocks, (p.layout.total_blocks + 255u) / 256u);
  }
  if (j1 > j0 && max_blocks) {
    const dim3 grid(std::min<uint32_t>(max_blocks, 4096u), (uint32_t)(j1 - j0));
    bgx_decondition_kernel<<<grid, 256, 0, st>>>(plan->d_precon + j0);
    BGX_CUDA(ctx, cudaGetLastError());
  }
  return bgx::kOk;
}

int bgx_plan_launch(bgx_context* ctx, bgx_plan* plan, void* cuda_stream) {
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
  plan->last_stream = st;
  plan->groups_used = 1;
  return launch_range(ctx, plan, 0, (uint32_t)plan->q_start.size() - 1, 0, st);
}

int bgx_plan_finish(bgx_context* ctx, bgx_plan* plan, uint32_t* bad_pages) {
  QueueCtl h[kMaxGroups];
  BGX_CUDA(ctx, cudaStreamSynchronize(plan->last_stream ? plan->last_stream : ctx->stream));
  BGX_CUDA(ctx, cudaMemcpy(h, plan->d_ctl, plan->groups_used * sizeof(QueueCtl), cudaMemcpyDeviceToHost));
  uint32_t bad = 0;
  for (uint32_t g = 0; g < plan->groups_used; ++g) bad += h[g].bad_pages;
  if (bad_pages) *bad_pages = bad;
  if (bad) { ctx->err = std::to_string(bad) + " page(s) failed to decode"; return bgx::kErrCorruptStream; }
  return bgx::kOk;
}

int bgx_decode_batch_host(bgx_context* ctx, uint32_t n, const uint8_t* const* inputs, const uint32_t* input_sizes,
                          uint8_t* const* outputs, uint32_t* output_sizes, double* kernel_ms) {
  BGX_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<StreamInfo> info(n);
  std::vector<size_t> in_off(n), out_off(n);
  size_t in_total = 0, out_total = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (input_sizes[i] < bgx::kStreamHeaderBytes) { ctx->err = "stream shorter than its header"; return bgx::kErrCorruptStream; }
    StreamInfo& si = info[i];
    const int rc = bgx::parse_stream_header(inputs[i], &si);
    if (rc) return rc;
    if (si.uncompressed_size > output_sizes[i]) { ctx->err = "output buffer too small"; return bgx::kErrGeneric; }
    in_off[i] = in_total;
    in_total += ((size_t)input_sizes[i] + bgx::kInputSlackBytes + 255u) & ~(size_t)255u;
    out_off[i] = out_total;
    out_total += ((size_t)si.uncompressed_size + 255u) & ~(size_t)255u;
  }
  if (grow(ctx, &ctx->d_in, &ctx->d_in_cap, in_total)) return bgx::kErrGeneric;
  if (grow(ctx, &ctx->d_out, &ctx->d_out_cap, out_total)) return bgx::kErrGeneric;

  // Three-stage pipeline: upload (s_in) -> kernels (ctx->stream) -> download (s_out); uploads and downloads run
  // concurrently on the two PCIe directions (the reference host serialises upload -> dispatch -> readback,
  // BrotligGPUDecoder.cpp:633-727). The unit of the pipeline is a segment (host_plan.h): a whole stream, or a page
  // range of a stream that is large against the batch, so a single 64 MiB stream overlaps its own three phases too.
  using Segment = bgx::HostSegment;
  const size_t total = in_total + out_total;
  const size_t target = std::max<size_t>(48u << 20, total / 12);   // (finer units measured no faster: PCIe is the bound)
  std::vector<Segment> seg;
  for (uint32_t i = 0; i < n; ++i) bgx::plan_stream_segments(i, inputs[i], input_sizes[i], info[i], target, seg);
  const uint32_t ns = (uint32_t)seg.size();
  std::vector<bgx_stream> st(ns);
  for (uint32_t k = 0; k < ns; ++k) {
    const Segment& g = seg[k];
    const uint32_t i = g.stream;
    bgx_stream& s = st[k];
    memset(&s, 0, sizeof s);
    s.d_src = ctx->d_in + in_off[i];
    s.src_size = input_sizes[i];
    s.src_capacity = (input_sizes[i] + 15u) & ~15u;   // the arena slot has kInputSlackBytes of slack
    s.d_dst = ctx->d_out + out_off[i] + g.dn0;        // where page `page_begin` goes
    s.dst_capacity = (uint32_t)(g.dn1 - g.dn0);
    s.page_begin = g.page_begin;
    s.page_count = g.page_count;
    memcpy(s.header, inputs[i], std::min<uint32_t>(16, input_sizes[i]));
  }
  bgx_plan* plan = nullptr;
  int rc = plan_create_impl(ctx, st.data(), ns, &plan, true);   // scratch planes from the context's arena
  if (rc) return rc;
  // groups of consecutive segments, each at least `target` bytes (one launch and one event triple per group)
  std::vector<uint32_t> cut{0};
  {
    size_t acc = 0;
    for (uint32_t k = 0; k < ns; ++k) {
      acc += (seg[k].up1 - seg[k].up0) + (seg[k].dn1 - seg[k].dn0);
      if ((acc >= target && cut.size() < kMaxGroups) || k + 1 == ns) { cut.push_back(k + 1); acc = 0; }
    }
  }
  const uint32_t G = (uint32_t)cut.size() - 1;
  plan->groups_used = G;
  // every CUDA call of the pipeline is checked: a failed copy must not turn into a "successful" decode of stale bytes
  struct PlanGuard { bgx_plan* p; ~PlanGuard() { bgx_plan_destroy(p); } } plan_guard{plan};
  while (ctx->ev_pool.size() < 3 * (size_t)G) {
    cudaEvent_t e;
    BGX_CUDA(ctx, cudaEventCreate(&e));
    ctx->ev_pool.push_back(e);
  }
  if (!ctx->ev_start) BGX_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
  const std::vector<cudaEvent_t>& ev = ctx->ev_pool;
  BGX_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));              // arena growth / earlier work on the context stream
  BGX_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_start, 0));
  BGX_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_start, 0));
  cudaError_t pipe_err = cudaSuccess;
  auto ok = [&](cudaError_t e) { if (e != cudaSuccess && pipe_err == cudaSuccess) pipe_err = e; return e == cudaSuccess; };
  for (uint32_t g = 0; g < G && !rc && pipe_err == cudaSuccess; ++g) {
    cudaEvent_t e_in = ev[3 * g], e_k0 = ev[3 * g + 1], e_k1 = ev[3 * g + 2];
    for (uint32_t k = cut[g]; k < cut[g + 1]; ++k) {
      const Segment& sg = seg[k];
      if (sg.up1 > sg.up0)
        ok(cudaMemcpyAsync(ctx->d_in + in_off[sg.stream] + sg.up0, inputs[sg.stream] + sg.up0, sg.up1 - sg.up0, cudaMemcpyHostToDevice, ctx->s_in));
    }
    ok(cudaEventRecord(e_in, ctx->s_in));
    ok(cudaStreamWaitEvent(ctx->stream, e_in, 0));
    ok(cudaEventRecord(e_k0, ctx->stream));
    if (pipe_err != cudaSuccess) break;
    rc = launch_range(ctx, plan, cut[g], cut[g + 1], g, ctx->stream);
    ok(cudaEventRecord(e_k1, ctx->stream));
    ok(cudaStreamWaitEvent(ctx->s_out, e_k1, 0));
    for (uint32_t k = cut[g]; k < cut[g + 1] && !rc; ++k) {
      const Segment& sg = seg[k];
      if (sg.dn1 > sg.dn0)
        ok(cudaMemcpyAsync(outputs[sg.stream] + sg.dn0, ctx->d_out + out_off[sg.stream] + sg.dn0, sg.dn1 - sg.dn0, cudaMemcpyDeviceToHost, ctx->s_out));
    }
  }
  // drain all three streams whatever happened (nothing may still be writing into the caller's buffers on return)
  ok(cudaStreamSynchronize(ctx->s_in));
  ok(cudaStreamSynchronize(ctx->stream));
  ok(cudaStreamSynchronize(ctx->s_out));
  if (pipe_err != cudaSuccess) return fail(ctx, "host-pointer pipeline", pipe_err);
  plan->last_stream = ctx->stream;
  if (!rc) rc = bgx_plan_finish(ctx, plan, nullptr);
  if (!rc) {
    double sum = 0;
    for (uint32_t g = 0; g < G; ++g) {
      float ms = 0;
      BGX_CUDA(ctx, cudaEventElapsedTime(&ms, ev[3 * g + 1], ev[3 * g + 2]));
      sum += ms;
    }
    if (kernel_ms) *kernel_ms += sum;
    for (uint32_t i = 0; i < n; ++i) output_sizes[i] = info[i].uncompressed_size;
  }
  return rc;
}

int bgx_decode_host_progress(bgx_context* ctx, uint32_t input_size, const uint8_t* input, uint32_t* output_size, uint8_t* output,
                             double* kernel_ms, bgx_progress_fn progress, void* user, uint32_t pages_per_group) {
  if (!progress) return bgx_decode_host(ctx, input_size, input, output_size, output, kernel_ms);
  BGX_CUDA(ctx, cudaSetDevice(ctx->device));
  if (input_size < bgx::kStreamHeaderBytes) { ctx->err = "stream shorter than its header"; return bgx::kErrCorruptStream; }
  StreamInfo si;
  int rc = bgx::parse_stream_header(input, &si);
  if (rc) return rc;
  if (si.uncompressed_size > *output_size) { ctx->err = "output buffer too small"; return bgx::kErrGeneric; }
  if (si.preconditioned) {
    // a page range of a texture scatters into the whole texture: one launch, then the per-page reports
    rc = bgx_decode_host(ctx, input_size, input, output_size, output, kernel_ms);
    for (uint32_t p = 0; p < si.num_pages && !rc; ++p)
      if (progress(user, p, si.num_pages)) break;
    return rc;
  }
  const size_t in_cap = ((size_t)input_size + bgx::kInputSlackBytes + 255u) & ~(size_t)255u;
  if (grow(ctx, &ctx->d_in, &ctx->d_in_cap, in_cap)) return bgx::kErrGeneric;
  if (grow(ctx, &ctx->d_out, &ctx->d_out_cap, ((size_t)si.uncompressed_size + 255u) & ~(size_t)255u)) return bgx::kErrGeneric;
  BGX_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, input, input_size, cudaMemcpyHostToDevice, ctx->stream));
  // groups of pages: the callback can stop the decode between two groups (BrotligDecoder.cpp:318-325 stops a worker
  // between two pages); the pages that were not decoded stay zero, as after the reference's memset (:448)
  uint32_t group = pages_per_group ? pages_per_group : std::max<uint32_t>(16u, (si.num_pages + 15u) / 16u);
  size_t done_bytes = 0;
  bool aborted = false;
  for (uint32_t b = 0; b < si.num_pages && !aborted; b += group) {
    const uint32_t c = std::min<uint32_t>(group, si.num_pages - b);
    size_t produced = (size_t)c * si.page_size;
    if (b + c == si.num_pages && si.last_page_size) produced -= si.page_size - si.last_page_size;
    bgx_stream s;
    memset(&s, 0, sizeof s);
    s.d_src = ctx->d_in;
    s.src_size = input_size;
    s.src_capacity = (input_size + 15u) & ~15u;
    s.d_dst = ctx->d_out + (size_t)b * si.page_size;
    s.dst_capacity = (uint32_t)produced;
    s.page_begin = b;
    s.page_count = c;
    memcpy(s.header, input, std::min<uint32_t>(16, input_size));
    bgx_plan* plan = nullptr;
    rc = plan_create_impl(ctx, &s, 1, &plan, true);
    if (rc) return rc;
    struct PlanGuard { bgx_plan* p; ~PlanGuard() { bgx_plan_destroy(p); } } plan_guard{plan};
    BGX_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    rc = launch_range(ctx, plan, 0, 1, 0, ctx->stream);
    if (rc) return rc;
    BGX_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    BGX_CUDA(ctx, cudaMemcpyAsync(output + (size_t)b * si.page_size, s.d_dst, produced, cudaMemcpyDeviceToHost, ctx->stream));
    plan->last_stream = ctx->stream;
    plan->groups_used = 1;
    rc = bgx_plan_finish(ctx, plan, nullptr);
    if (rc) return rc;
    float ms = 0;
    BGX_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (kernel_ms) *kernel_ms += ms;
    done_bytes = (size_t)b * si.page_size + produced;
    for (uint32_t p = b; p < b + c; ++p)
      if (progress(user, p, si.num_pages)) { aborted = true; break; }
  }
  if (aborted && done_bytes < si.uncompressed_size) memset(output + done_bytes, 0, si.uncompressed_size - done_bytes);
  *output_size = si.uncompressed_size;
  return bgx::kOk;
}

// One process, several devices: the streams of the batch are split over the contexts (size balanced, whole streams,
// no inter-GPU traffic -- pages and streams are independent) and every context decodes its share on its own host thread.
int bgx_decode_batch_host_multi(bgx_context* const* ctxs, uint32_t n_ctx, uint32_t n, const uint8_t* const* inputs,
                                const uint32_t* input_sizes, uint8_t* const* outputs, uint32_t* output_sizes, double* kernel_ms) {
  if (n_ctx == 0) return bgx::kErrGeneric;
  if (n_ctx == 1) return bgx_decode_batch_host(ctxs[0], n, inputs, input_sizes, outputs, output_sizes, kernel_ms);
  // longest-processing-time-first by compressed size
  std::vector<uint32_t> order(n);
  for (uint32_t i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return input_sizes[a] != input_sizes[b] ? input_sizes[a] > input_sizes[b] : a < b; });
  std::vector<std::vector<uint32_t>> share(n_ctx);
  std::vector<uint64_t> load(n_ctx, 0);
  for (uint32_t i : order) {
    const uint32_t r = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
    share[r].push_back(i);
    load[r] += input_sizes[i];
  }
  std::vector<int> rcs(n_ctx, 0);
  std::vector<double> ms(n_ctx, 0.0);
  std::vector<std::thread> th;
  for (uint32_t r = 0; r < n_ctx; ++r) {
    th.emplace_back([&, r] {
      const std::vector<uint32_t>& mine = share[r];
      if (mine.empty()) return;
      std::vector<const uint8_t*> in(mine.size());
      std::vector<uint32_t> isz(mine.size()), osz(mine.size());
      std::vector<uint8_t*> out(mine.size());
      for (size_t k = 0; k < mine.size(); ++k) { in[k] = inputs[mine[k]]; isz[k] = input_sizes[mine[k]]; out[k] = outputs[mine[k]]; osz[k] = output_sizes[mine[k]]; }
      rcs[r] = bgx_decode_batch_host(ctxs[r], (uint32_t)mine.size(), in.data(), isz.data(), out.data(), osz.data(), &ms[r]);
      for (size_t k = 0; k < mine.size(); ++k) output_sizes[mine[k]] = osz[k];
    });
  }
  for (auto& t : th) t.join();
  int rc = 0;
  double worst = 0;
  for (uint32_t r = 0; r < n_ctx; ++r) {
    if (rcs[r] && !rc) rc = rcs[r];
    worst = std::max(worst, ms[r]);
  }
  if (kernel_ms) *kernel_ms += worst;   // devices run concurrently: the slowest one is the batch's kernel time
  return rc;
}

int bgx_decode_host(bgx_context* ctx, uint32_t input_size, const uint8_t* input, uint32_t* output_size, uint8_t* output,
                    double* kernel_ms) {
  const uint8_t* ins[1] = {input};
  uint8_t* outs[1] = {output};
  return bgx_decode_batch_host(ctx, 1, ins, &input_size, outs, output_size, kernel_ms);
}

}  // extern "C"

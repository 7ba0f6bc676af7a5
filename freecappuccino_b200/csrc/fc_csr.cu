// CSR pattern and cell-to-face map, built on the device.
//
// Reference: create_CSR_matrix_from_mesh_data (src/sparse_matrix.f90:42-172) builds
// COO = (i,i) ++ (owner,neigh) ++ (neigh,owner), heap-sorts it lexicographically
// (src/utils.f90:424-770) and derives ioffset / diag / the two face->CSR maps by
// linear search (csr_to_k, src/utils.f90:76-100).  The sorted pattern is unique, so
// here it is produced by count -> scan -> fill -> per-row sort, which gives
// bit-identical integers without a global sort.
#include <vector>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_reduce.cuh>

#include "fc_internal.cuh"

namespace {

__global__ void k_fill_int(int *p, int v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void k_count_inner(const int *__restrict__ owner, const int *__restrict__ neigh, int F, int *cnt) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  atomicAdd(&cnt[owner[f]], 1);
  atomicAdd(&cnt[neigh[f]], 1);
}

// row c starts with its own diagonal entry; `fill` = next free position
__global__ void k_rows_init(const int *__restrict__ ioffset, int n, int *ja, int *fill) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  int s = ioffset[c];
  ja[s] = c;
  fill[c] = s + 1;
}

__global__ void k_rows_fill(const int *__restrict__ owner, const int *__restrict__ neigh, int F, int *fill,
                            int *ja) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  int p = owner[f], q = neigh[f];
  ja[atomicAdd(&fill[p], 1)] = q;
  ja[atomicAdd(&fill[q], 1)] = p;
}

// ascending insertion sort of every (short) row; diag = first position holding the row id
__global__ void k_rows_sort(const int *__restrict__ ioffset, int n, int *ja, int *diag, int *dupflag) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  int s = ioffset[c], e = ioffset[c + 1];
  for (int i = s + 1; i < e; ++i) {
    int v = ja[i], j = i - 1;
    while (j >= s && ja[j] > v) { ja[j + 1] = ja[j]; --j; }
    ja[j + 1] = v;
  }
  int d = -1;
  for (int i = s; i < e; ++i) {
    if (ja[i] == c && d < 0) d = i;
    if (i > s && ja[i] == ja[i - 1]) *dupflag = 1;
  }
  diag[c] = d;
}

__device__ __forceinline__ int row_find(const int *__restrict__ ioffset, const int *__restrict__ ja, int row,
                                        int col) {
  for (int k = ioffset[row]; k < ioffset[row + 1]; ++k)
    if (ja[k] == col) return k;  // first match, like csr_to_k
  return -1;
}

__global__ void k_face_maps(const int *__restrict__ owner, const int *__restrict__ neigh, int F,
                            const int *__restrict__ ioffset, const int *__restrict__ ja, int *icj, int *jci) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  int p = owner[f], q = neigh[f];
  icj[f] = row_find(ioffset, ja, p, q);
  jci[f] = row_find(ioffset, ja, q, p);
}

// position of the transposed entry a(j,i) for every strictly-lower a(i,j)
// (the search of bicgstab.f90:72-75 runs from diag(j) to the row end)
__global__ void k_transpose_pos(const int *__restrict__ ioffset, const int *__restrict__ ja,
                                const int *__restrict__ diag, int n, int *tpos) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = ioffset[i]; k < diag[i]; ++k) {
    int j = ja[k], t = -1;
    for (int l = diag[j]; l < ioffset[j + 1]; ++l)
      if (ja[l] == i) { t = l; break; }
    tpos[k] = t;
  }
}

__global__ void k_chunk_nnz(const int *__restrict__ ioffset, int n, int rows, int *out, int nchunks) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nchunks) return;
  int r0 = b * rows, r1 = min(n, r0 + rows);
  out[b] = ioffset[r1] - ioffset[r0];
}

// ---- cell -> face map -------------------------------------------------------
struct kinds_t {
  int fstart[6];  // 0-based first face of: proc, inlet, outlet, symmetry, wall, prOutlet
  int count[6];
  int slot[6];    // 0-based first field slot of each
};

__device__ __forceinline__ int boundary_slot(const kinds_t &K, int f) {
  for (int b = 0; b < 6; ++b)
    if (K.count[b] > 0 && f >= K.fstart[b] && f < K.fstart[b] + K.count[b]) return K.slot[b] + (f - K.fstart[b]);
  return -1;
}

__global__ void k_c2f_count(const int *__restrict__ owner, const int *__restrict__ neigh, int F, int NF,
                            kinds_t K, int *cnt) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= NF) return;
  if (f < F) {
    atomicAdd(&cnt[owner[f]], 1);
    atomicAdd(&cnt[neigh[f]], 1);
  } else if (boundary_slot(K, f) >= 0) {
    atomicAdd(&cnt[owner[f]], 1);
  }
}

__global__ void k_c2f_fill(const int *__restrict__ owner, const int *__restrict__ neigh, int F, int NF, kinds_t K,
                           int *fill, int *face, int *other) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= NF) return;
  if (f < F) {
    int p = owner[f], q = neigh[f];
    int e = atomicAdd(&fill[p], 1);
    face[e] = f;
    other[e] = q;
    e = atomicAdd(&fill[q], 1);
    face[e] = f | 0x80000000;
    other[e] = p;
  } else {
    int s = boundary_slot(K, f);
    if (s < 0) return;
    int e = atomicAdd(&fill[owner[f]], 1);
    face[e] = f;
    other[e] = s;
  }
}

// sort the entries of every cell into the reference's loop order and attach the CSR position
__global__ void k_c2f_sort(const int *__restrict__ off, int n, int F, int *face, int *other, int *pos,
                           const int *__restrict__ ioffset, const int *__restrict__ ja) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  int s = off[c], e = off[c + 1];
  auto key = [&](int i) {
    int f = face[i] & 0x7fffffff;
    return f < F ? f : F + (other[i] - n);  // boundary: slot order = proc, inlet, outlet, symmetry, wall, prOutlet
  };
  for (int i = s + 1; i < e; ++i) {
    int kf = face[i], ko = other[i], kk = key(i), j = i - 1;
    while (j >= s && key(j) > kk) { face[j + 1] = face[j]; other[j + 1] = other[j]; --j; }
    face[j + 1] = kf;
    other[j + 1] = ko;
  }
  for (int i = s; i < e; ++i) {
    int f = face[i] & 0x7fffffff;
    pos[i] = (f < F && ioffset) ? row_find(ioffset, ja, c, other[i]) : -1;
  }
}

// per-row list of processor faces (ascending i): the `apr` strip the SpMV adds after the CSR part
// of a row (src-parallel/dpcg.f90:132-136)
__global__ void k_strip_count(const int *__restrict__ off, const int *__restrict__ other, int n, int npro, int F,
                              const int *__restrict__ face, int *cnt) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  int k = 0;
  for (int q = off[c]; q < off[c + 1]; ++q)
    if ((face[q] & 0x7fffffff) >= F && other[q] < n + npro) ++k;
  cnt[c] = k;
}
__global__ void k_strip_fill(const int *__restrict__ off, const int *__restrict__ other, int n, int npro, int F,
                             const int *__restrict__ face, const int *__restrict__ soff, int *sidx) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  int k = soff[c];
  for (int q = off[c]; q < off[c + 1]; ++q)
    if ((face[q] & 0x7fffffff) >= F && other[q] < n + npro) sidx[k++] = other[q] - n;
}

__global__ void k_strip_any32(const int *__restrict__ soff, int n, unsigned char *any32) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g * 32 >= n) return;
  int r1 = min(n, (g + 1) * 32);
  any32[g] = soff[r1] > soff[g * 32] ? 1 : 0;
}

// ---- column indices as one-byte codes (CODED pipelines of fc_spmv_pipe.cuh) ----
// bit (ja[k] - row + n) of `bits` for every non-zero; the bit is tested before the atomic, so the few distinct
// offsets of a finite-volume numbering cost a handful of atomics, not one per non-zero
__global__ void k_delta_mark(const int *__restrict__ ioffset, const int *__restrict__ ja, int n, unsigned *bits) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (int k = ioffset[r]; k < ioffset[r + 1]; ++k) {
    const unsigned d = (unsigned)(ja[k] - r + n);
    const unsigned m = 1u << (d & 31u);
    if (!(__ldcg(bits + (d >> 5)) & m)) atomicOr(bits + (d >> 5), m);
  }
}
// code[k] = position of (ja[k] - row) in the ascending dictionary
__global__ void k_delta_code(const int *__restrict__ ioffset, const int *__restrict__ ja, int n,
                             const int *__restrict__ dict, int nd, unsigned char *code) {
  __shared__ int s_d[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_d[i] = dict[i];
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (int k = ioffset[r]; k < ioffset[r + 1]; ++k) {
    const int d = ja[k] - r;
    int lo = 0, hi = nd - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_d[mid] < d) lo = mid + 1;
      else hi = mid;
    }
    code[k] = (unsigned char)lo;
  }
}

int exclusive_scan(fc_context *ctx, int *in, int *out, int count) {
  void *tmp = nullptr;
  size_t bytes = 0;
  FC_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, count, ctx->stream));
  FC_CUDA(cudaMalloc(&tmp, bytes ? bytes : 1));
  cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, count, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(tmp);
  FC_CUDA(e);
  ctx->launches++;
  return FC_OK;
}

}  // namespace

int fc_csr_build(fc_context *ctx) {
  const int n = ctx->n, F = ctx->F, nnz = ctx->nnz;
  const int B = 256;
  int *cnt = nullptr, *fill = nullptr, *dup = nullptr;
  FC_CHECK(fc_dev_alloc(ctx, &cnt, (size_t)n + 1));
  FC_CHECK(fc_dev_alloc(ctx, &fill, (size_t)n));
  FC_CHECK(fc_dev_alloc(ctx, &dup, 1));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->ioffset, (size_t)n + 1));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->ja, (size_t)nnz));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->diag, (size_t)n));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->icj, (size_t)F));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->jci, (size_t)F));
  FC_CUDA(cudaMemsetAsync(dup, 0, sizeof(int), ctx->stream));
  k_fill_int<<<fc_blocks(n + 1, B), B, 0, ctx->stream>>>(cnt, 1, n + 1);
  FC_LAUNCH_CHECK();
  if (F > 0) {
    k_count_inner<<<fc_blocks(F, B), B, 0, ctx->stream>>>(ctx->owner, ctx->neigh, F, cnt);
    FC_LAUNCH_CHECK();
  }
  FC_CHECK(exclusive_scan(ctx, cnt, ctx->ioffset, n + 1));
  k_rows_init<<<fc_blocks(n, B), B, 0, ctx->stream>>>(ctx->ioffset, n, ctx->ja, fill);
  FC_LAUNCH_CHECK();
  if (F > 0) {
    k_rows_fill<<<fc_blocks(F, B), B, 0, ctx->stream>>>(ctx->owner, ctx->neigh, F, fill, ctx->ja);
    FC_LAUNCH_CHECK();
  }
  k_rows_sort<<<fc_blocks(n, B), B, 0, ctx->stream>>>(ctx->ioffset, n, ctx->ja, ctx->diag, dup);
  FC_LAUNCH_CHECK();
  if (F > 0) {
    k_face_maps<<<fc_blocks(F, B), B, 0, ctx->stream>>>(ctx->owner, ctx->neigh, F, ctx->ioffset, ctx->ja, ctx->icj,
                                                        ctx->jci);
    FC_LAUNCH_CHECK();
  }
  int hdup = 0;
  FC_CUDA(cudaMemcpyAsync(&hdup, dup, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->csr_dup = hdup != 0;
  cudaFree(cnt);
  cudaFree(fill);
  cudaFree(dup);
  return fc_csr_post(ctx);
}

// derived data every solver needs: transposed positions (DILU) and the largest
// nnz count of a 256-row block (selects the SpMV staging size)
int fc_csr_post(fc_context *ctx) {
  const int n = ctx->n, B = 256;
  FC_CHECK(fc_dev_alloc(ctx, &ctx->tpos, (size_t)ctx->nnz));
  k_transpose_pos<<<fc_blocks(n, B), B, 0, ctx->stream>>>(ctx->ioffset, ctx->ja, ctx->diag, n, ctx->tpos);
  FC_LAUNCH_CHECK();
  const int rows = 256, nchunks = (n + rows - 1) / rows;
  int *chunk = nullptr, *mx = nullptr;
  FC_CHECK(fc_dev_alloc(ctx, &chunk, (size_t)nchunks));
  FC_CHECK(fc_dev_alloc(ctx, &mx, 1));
  k_chunk_nnz<<<fc_blocks(nchunks, B), B, 0, ctx->stream>>>(ctx->ioffset, n, rows, chunk, nchunks);
  FC_LAUNCH_CHECK();
  void *tmp = nullptr;
  size_t bytes = 0;
  FC_CUDA(cub::DeviceReduce::Max(tmp, bytes, chunk, mx, nchunks, ctx->stream));
  FC_CUDA(cudaMalloc(&tmp, bytes ? bytes : 1));
  cudaError_t e = cub::DeviceReduce::Max(tmp, bytes, chunk, mx, nchunks, ctx->stream);
  int h = 0;
  cudaMemcpyAsync(&h, mx, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(tmp);
  cudaFree(chunk);
  cudaFree(mx);
  FC_CUDA(e);
  ctx->launches++;
  ctx->spmv_max_chunk = h;
  ctx->has_csr = true;
  return fc_codes_build(ctx);
}

// One-byte column codes of the current pattern: ja[k] = row + jdict[jcode[k]].  Possible when the pattern has at most
// 256 distinct column offsets (structured and block-structured numberings; 7 on a hexahedral box); otherwise
// coded_ok stays false and every kernel reads `ja`.
int fc_codes_build(fc_context *ctx) {
  const int n = ctx->n, B = 256;
  ctx->coded_ok = false;
  if (n < 1 || ctx->nnz < 1) return FC_OK;
  const size_t words = ((size_t)2 * n + 1 + 31) / 32;
  unsigned *bits = nullptr;
  FC_CHECK(fc_dev_alloc(ctx, &bits, words));
  FC_CUDA(cudaMemsetAsync(bits, 0, words * sizeof(unsigned), ctx->stream));
  k_delta_mark<<<fc_blocks(n, B), B, 0, ctx->stream>>>(ctx->ioffset, ctx->ja, n, bits);
  FC_LAUNCH_CHECK();
  std::vector<unsigned> hb(words);
  FC_CUDA(cudaMemcpyAsync(hb.data(), bits, words * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(bits);
  std::vector<int> dict;
  for (size_t w = 0; w < words && dict.size() <= 256; ++w) {
    unsigned x = hb[w];
    while (x && dict.size() <= 256) {
      const int b = __builtin_ctz(x);
      x &= x - 1;
      dict.push_back((int)((long long)w * 32 + b - n));
    }
  }
  if (dict.empty() || dict.size() > 256) return FC_OK;
  const int nd = (int)dict.size();
  dict.resize(256, dict.back());
  FC_CHECK(fc_dev_alloc(ctx, &ctx->jdict, 256));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->jcode, (size_t)ctx->nnz + 64));
  FC_CUDA(cudaMemcpyAsync(ctx->jdict, dict.data(), 256 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  k_delta_code<<<fc_blocks(n, B), B, 0, ctx->stream>>>(ctx->ioffset, ctx->ja, n, ctx->jdict, nd, ctx->jcode);
  FC_LAUNCH_CHECK();
  FC_CUDA(cudaStreamSynchronize(ctx->stream));   // `dict` is a host temporary
  ctx->launches += 2;
  ctx->ndict = nd;
  ctx->coded_ok = true;
  return FC_OK;
}

int fc_c2f_build(fc_context *ctx) {
  const int n = ctx->n, F = ctx->F, NF = ctx->NF, B = 256;
  const fc_mesh_desc &m = ctx->m;
  kinds_t K;
  const int cnts[6] = {m.npro, m.ninl, m.nout, m.nsym, m.nwal, m.npru};
  const int fst[6] = {m.iProcFacesStart, m.iInletFacesStart, m.iOutletFacesStart,
                      m.iSymmetryFacesStart, m.iWallFacesStart, m.iPressOutletFacesStart};
  int slot = n;
  for (int b = 0; b < 6; ++b) {
    K.count[b] = cnts[b];
    K.fstart[b] = fst[b];
    K.slot[b] = slot;
    slot += cnts[b];
  }
  int *cnt = nullptr, *fill = nullptr;
  FC_CHECK(fc_dev_alloc(ctx, &cnt, (size_t)n + 1));
  FC_CHECK(fc_dev_alloc(ctx, &fill, (size_t)n + 1));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->c2f_off, (size_t)n + 1));
  FC_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)n + 1), ctx->stream));
  k_c2f_count<<<fc_blocks(NF, B), B, 0, ctx->stream>>>(ctx->owner, ctx->neigh, F, NF, K, cnt);
  FC_LAUNCH_CHECK();
  FC_CHECK(exclusive_scan(ctx, cnt, ctx->c2f_off, n + 1));
  int len = 0;
  FC_CUDA(cudaMemcpy(&len, ctx->c2f_off + n, sizeof(int), cudaMemcpyDeviceToHost));
  ctx->c2f_len = len;
  FC_CHECK(fc_dev_alloc(ctx, &ctx->c2f_face, (size_t)len));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->c2f_other, (size_t)len));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->c2f_pos, (size_t)len));
  FC_CUDA(cudaMemcpyAsync(fill, ctx->c2f_off, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, ctx->stream));
  k_c2f_fill<<<fc_blocks(NF, B), B, 0, ctx->stream>>>(ctx->owner, ctx->neigh, F, NF, K, fill, ctx->c2f_face,
                                                      ctx->c2f_other);
  FC_LAUNCH_CHECK();
  k_c2f_sort<<<fc_blocks(n, B), B, 0, ctx->stream>>>(ctx->c2f_off, n, F, ctx->c2f_face, ctx->c2f_other,
                                                     ctx->c2f_pos, ctx->has_csr ? ctx->ioffset : nullptr, ctx->ja);
  FC_LAUNCH_CHECK();
  if (ctx->npro > 0) {
    FC_CHECK(fc_dev_alloc(ctx, &ctx->strip_off, (size_t)n + 1));
    FC_CHECK(fc_dev_alloc(ctx, &ctx->strip_idx, (size_t)ctx->npro));
    FC_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)n + 1), ctx->stream));
    k_strip_count<<<fc_blocks(n, B), B, 0, ctx->stream>>>(ctx->c2f_off, ctx->c2f_other, n, ctx->npro, F, ctx->c2f_face,
                                                          cnt);
    FC_LAUNCH_CHECK();
    FC_CHECK(exclusive_scan(ctx, cnt, ctx->strip_off, n + 1));
    k_strip_fill<<<fc_blocks(n, B), B, 0, ctx->stream>>>(ctx->c2f_off, ctx->c2f_other, n, ctx->npro, F, ctx->c2f_face,
                                                         ctx->strip_off, ctx->strip_idx);
    FC_LAUNCH_CHECK();
    FC_CHECK(fc_dev_alloc(ctx, &ctx->strip_any32, (size_t)(n + 31) / 32));
    k_strip_any32<<<fc_blocks((size_t)(n + 31) / 32, B), B, 0, ctx->stream>>>(ctx->strip_off, n, ctx->strip_any32);
    FC_LAUNCH_CHECK();
  }
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(cnt);
  cudaFree(fill);
  return FC_OK;
}

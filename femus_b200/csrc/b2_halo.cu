// Distributed layout of rank-local vectors (one rank per GPU, mesh sharded by element).
// Replaces what PETSc does behind PetscVector / PetscMatrix for an MPI run of the reference:
//   - VecSetValues(ADD_VALUES) + VecAssemblyBegin/End: off-process contributions summed into the
//     owner (PetscVector.cpp:132-141, PetscVector.hpp:595-602),
//   - VecGhostUpdateBegin/End inside close() (PetscVector.hpp:604-609) and the VecScatter inside
//     MatMult: ghost values refreshed from the owner,
//   - VecDot / VecNorm over the owned entries + MPI_Allreduce (PetscVector.cpp:43-87, 399-410).
// Here every rank keeps ALL dofs of its own elements (its matrices are the sums over its own
// elements only), so "assemble off-process contributions" and "refresh ghosts" collapse into ONE
// operation: sum the interface entries over the ranks that hold them.  The interface dofs of all
// ranks are laid out in one packed vector (position = rank-independent, computed by the host from
// the lattice names of the nodes); each rank writes its partial values at its own positions of a
// send buffer that is zero elsewhere, one ncclAllReduce(sum) over NVLink/NVSwitch completes them,
// and the rank reads back its own positions.
#include "b2_common.cuh"
#include "b2_peer.cuh"

namespace {
constexpr int kBlock = 256;

__global__ void halo_pack_kernel(int64_t n, const int32_t* __restrict__ idx, const int32_t* __restrict__ pos,
                                 const double* __restrict__ v, double* __restrict__ send) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) send[pos[k]] = v[idx[k]];
}
__global__ void halo_unpack_kernel(int64_t n, const int32_t* __restrict__ idx, const int32_t* __restrict__ pos,
                                   const double* __restrict__ recv, double* __restrict__ v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) v[idx[k]] = recv[pos[k]];
}
// the same with nscal extra scalars appended to the packed vector (one collective for both)
__global__ void halo_pack_scal_kernel(int64_t n, const int32_t* __restrict__ idx, const int32_t* __restrict__ pos,
                                      const double* __restrict__ v, double* __restrict__ send, int64_t n_packed,
                                      const double* __restrict__ scal, int nscal) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t0 < nscal) send[n_packed + t0] = scal[t0];
  for (int64_t k = t0; k < n; k += stride) send[pos[k]] = v[idx[k]];
}
__global__ void halo_unpack_scal_kernel(int64_t n, const int32_t* __restrict__ idx, const int32_t* __restrict__ pos,
                                        const double* __restrict__ recv, double* __restrict__ v, int64_t n_packed,
                                        double* __restrict__ scal, int nscal) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t0 < nscal) scal[t0] = recv[n_packed + t0];
  for (int64_t k = t0; k < n; k += stride) v[idx[k]] = recv[pos[k]];
}
__global__ void halo_invmult_kernel(int64_t n, const uint8_t* __restrict__ mult, double* __restrict__ inv) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) inv[i] = 1.0 / (double)mult[i];
}

// ---- peer-memory exchange ------------------------------------------------------------------------------------
// One interface sum = ONE kernel, no NCCL call, no fence: the thread that owns an interface entry stores its value
// straight into the inbox of every other rank holding that dof (remote stores over NVLink / NVSwitch), then waits for
// their values in its own inbox and sums over the holders in ascending rank order (own value in place) -- all holders
// of a dof compute the same bits.  A value travels as a 16-byte cell {low word, flag, high word, flag} with flag = the
// exchange number (the "LL" protocol of NCCL: 8-byte stores are atomic, so a half is either old or complete, and the
// receiver simply re-reads until both flags match).  Scalars (dot products of the coarse solver) ride in the first 8
// cells of every slot, from every rank to every rank, summed in rank order.  Slots are double buffered by the parity
// of the exchange number: a rank cannot finish exchange e + 1 before every sharing rank has finished e.
__global__ void peer_ll_kernel(void* const* __restrict__ bases, void* base, int nranks, int me, int64_t slot, unsigned long long epoch,
                               int64_t n_if, const int32_t* __restrict__ idx, const int64_t* __restrict__ hold_ptr,
                               const int32_t* __restrict__ hold_rank, const int32_t* __restrict__ hold_pos,
                               const int32_t* __restrict__ hold_spos, double* __restrict__ v, double* __restrict__ scal, int nscal,
                               int* __restrict__ err) {
  const int parity = (int)(epoch & 1ull);
  const unsigned flag = b2_peer_flag(epoch);
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (blockIdx.x == 0 && threadIdx.x < nscal) {
    const double mine = scal[threadIdx.x];
    for (int r = 0; r < nranks; r++)
      if (r != me) b2_peer_store(b2_peer_cell(bases[r], slot, nranks, parity, me, threadIdx.x), mine, flag);
  }
  for (int64_t k = t0; k < n_if; k += stride) {
    const int32_t d = idx[k];
    const double own = v[d];
    const int64_t h0 = hold_ptr[k], h1 = hold_ptr[k + 1];
    for (int64_t q = h0; q < h1; q++)
      if (hold_rank[q] != me) b2_peer_store(b2_peer_cell(bases[hold_rank[q]], slot, nranks, parity, me, kPeerScal + hold_spos[q]), own, flag);
    double s = 0.0;
    for (int64_t q = h0; q < h1; q++) {
      const int r = hold_rank[q];
      const double a = r == me ? own : b2_peer_wait(b2_peer_cell(base, slot, nranks, parity, r, kPeerScal + hold_pos[q]), flag, err);
      s = q == h0 ? a : s + a;
    }
    v[d] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x < nscal) {
    const double mine = scal[threadIdx.x];
    double s = 0.0;
    for (int r = 0; r < nranks; r++) {
      const double a = r == me ? mine : b2_peer_wait(b2_peer_cell(base, slot, nranks, parity, r, threadIdx.x), flag, err);
      s = r == 0 ? a : s + a;
    }
    scal[threadIdx.x] = s;
  }
}

bool use_peer(const b2_halo* h) { return h->ctx->halo_peer && h->ctx->peer_base && h->hold_ptr; }

int peer_exchange(b2_halo* h, b2_vec* v, double* d_scal, int nscal) {
  b2_ctx* c = h->ctx;
  const unsigned long long epoch = ++c->peer_epoch;
  const int grid = b2_grid_for(c, h->n_if > 0 ? h->n_if : 1, kBlock, 4);
  B2_LAUNCH(c, peer_ll_kernel, grid, kBlock, 0, (void* const*)c->d_peer_base, c->peer_local, c->nranks, c->rank, c->peer_slot, epoch, h->n_if,
            h->idx, h->hold_ptr, h->hold_rank, h->hold_pos, h->hold_spos, v->d, d_scal, nscal, c->peer_err);
  return 0;
}

}  // namespace

int b2_allreduce_into(b2_ctx* c, const double* d_send, double* d_recv, int64_t n);   // b2_ctx.cu

extern "C" {

int b2_halo_create(b2_ctx* c, int64_t n_local, int64_t n_if, const int32_t* local_idx, const int32_t* packed_pos,
                   int64_t n_packed, const uint8_t* owned, const uint8_t* mult, b2_halo** out) {
  *out = nullptr;
  B2_CHECK(c && n_local >= 0 && n_if >= 0 && n_packed >= n_if && owned && mult, "b2_halo_create: bad arguments");
  B2_CHECK(n_if == 0 || (local_idx && packed_pos), "b2_halo_create: null index arrays");
  for (int64_t k = 0; k < n_if; k++)
    B2_CHECK(local_idx[k] >= 0 && local_idx[k] < n_local && packed_pos[k] >= 0 && packed_pos[k] < n_packed,
             "b2_halo_create: interface entry %lld out of range", (long long)k);
  b2_halo* h = new b2_halo();
  h->ctx = c;
  h->n_local = n_local;
  h->n_if = n_if;
  h->n_packed = n_packed;
  h->n_owned = 0;
  for (int64_t i = 0; i < n_local; i++) {
    B2_CHECK(mult[i] >= 1, "b2_halo_create: multiplicity of dof %lld is zero", (long long)i);
    h->n_owned += owned[i] ? 1 : 0;
  }
  B2_TRY(b2_malloc(c, &h->idx, (size_t)n_if));
  B2_TRY(b2_malloc(c, &h->pos, (size_t)n_if));
  B2_TRY(b2_malloc(c, &h->send, (size_t)n_packed + 8));     // +8: scalars riding with the interface sum
  B2_TRY(b2_malloc(c, &h->recv, (size_t)n_packed + 8));
  B2_TRY(b2_malloc(c, &h->owned, (size_t)n_local));
  B2_TRY(b2_malloc(c, &h->invmult, (size_t)n_local + 4));   // +4: the SpMV stages 16-byte aligned slices
  uint8_t* d_mult = nullptr;
  B2_TRY(b2_malloc(c, &d_mult, (size_t)n_local));
  B2_TRY(b2_upload(c, h->idx, local_idx, (size_t)n_if));
  B2_TRY(b2_upload(c, h->pos, packed_pos, (size_t)n_if));
  B2_TRY(b2_upload(c, h->owned, owned, (size_t)n_local));
  B2_TRY(b2_upload(c, d_mult, mult, (size_t)n_local));
  B2_CUDA(cudaMemsetAsync(h->send, 0, (size_t)(n_packed + 8) * sizeof(double), c->stream));
  B2_CUDA(cudaMemsetAsync(h->recv, 0, (size_t)(n_packed + 8) * sizeof(double), c->stream));
  if (n_local) B2_LAUNCH(c, halo_invmult_kernel, b2_grid_for(c, n_local, kBlock, 8), kBlock, 0, n_local, d_mult, h->invmult);
  B2_CUDA(cudaStreamSynchronize(c->stream));
  b2_free(c, d_mult, (size_t)n_local);
  *out = h;
  return 0;
}

int b2_halo_destroy(b2_halo* h) {
  if (!h) return 0;
  b2_ctx* c = h->ctx;
  cudaStreamSynchronize(c->stream);
  b2_free(c, h->idx, (size_t)h->n_if);
  b2_free(c, h->pos, (size_t)h->n_if);
  b2_free(c, h->send, (size_t)h->n_packed + 8);
  b2_free(c, h->recv, (size_t)h->n_packed + 8);
  b2_free(c, h->owned, (size_t)h->n_local);
  b2_free(c, h->invmult, (size_t)h->n_local + 4);
  if (h->hold_ptr) {
    b2_free(c, h->hold_ptr, (size_t)h->n_if + 1);
    b2_free(c, h->hold_rank, (size_t)h->n_hold);
    b2_free(c, h->hold_pos, (size_t)h->n_hold);
    b2_free(c, h->hold_spos, (size_t)h->n_hold);
  }
  delete h;
  return 0;
}

int64_t b2_halo_owned_count(const b2_halo* h) { return h->n_owned; }
int64_t b2_halo_interface_count(const b2_halo* h) { return h->n_if; }

/* v[interface] <- sum over the ranks holding each interface dof of their v[interface] */
int b2_halo_sum(b2_halo* h, b2_vec* v) {
  B2_CHECK(v->n >= h->n_local, "b2_halo_sum: vector shorter than the layout");
  b2_ctx* c = h->ctx;
  if (c->nranks == 1 || h->n_packed == 0) return 0;
  if (use_peer(h)) return peer_exchange(h, v, nullptr, 0);
  if (h->n_if) B2_LAUNCH(c, halo_pack_kernel, b2_grid_for(c, h->n_if, kBlock, 8), kBlock, 0, h->n_if, h->idx, h->pos, v->d, h->send);
  B2_TRY(b2_allreduce_into(c, h->send, h->recv, h->n_packed));
  if (h->n_if) B2_LAUNCH(c, halo_unpack_kernel, b2_grid_for(c, h->n_if, kBlock, 8), kBlock, 0, h->n_if, h->idx, h->pos, h->recv, v->d);
  return 0;
}

}  // extern "C"

// v[interface] <- sum over ranks, and d_scal[0..nscal) <- sum over ranks, in ONE ncclAllReduce
// (latency-bound solvers: the coarse PCG sends its dot products with the interface values)
int b2_halo_sum_scalars(b2_halo* h, b2_vec* v, double* d_scal, int nscal) {
  B2_CHECK(v->n >= h->n_local && nscal >= 0 && nscal <= 8, "b2_halo_sum_scalars: bad arguments");
  b2_ctx* c = h->ctx;
  if (c->nranks == 1) return 0;
  if (use_peer(h)) return peer_exchange(h, v, d_scal, nscal);
  const int grid = b2_grid_for(c, h->n_if > 0 ? h->n_if : 1, kBlock, 8);
  B2_LAUNCH(c, halo_pack_scal_kernel, grid, kBlock, 0, h->n_if, h->idx, h->pos, v->d, h->send, h->n_packed, d_scal, nscal);
  B2_TRY(b2_allreduce_into(c, h->send, h->recv, h->n_packed + nscal));
  B2_LAUNCH(c, halo_unpack_scal_kernel, grid, kBlock, 0, h->n_if, h->idx, h->pos, h->recv, v->d, h->n_packed, d_scal, nscal);
  return 0;
}

extern "C" {

/* Peer-memory form of the interface sum (needs b2_ctx_peer_open): for every interface entry k (order of local_idx at
 * creation) its holders in ascending rank order, this rank included; for each holder the cell of ITS value in its
 * message to this rank (hold_pos) and the cell of THIS rank's value in its message to the holder (hold_spos).  A
 * message rank q -> rank r lists, in q's entry order, the entries of q that r also holds. */
int b2_halo_set_exchange(b2_halo* h, const int64_t* hold_ptr, const int32_t* hold_rank, const int32_t* hold_pos, const int32_t* hold_spos) {
  B2_CHECK(h && hold_ptr && (h->n_if == 0 || (hold_rank && hold_pos && hold_spos)), "b2_halo_set_exchange: bad arguments");
  B2_CHECK(!h->hold_ptr, "b2_halo_set_exchange: already set");
  b2_ctx* c = h->ctx;
  const int64_t nh = hold_ptr[h->n_if];
  for (int64_t q = 0; q < nh; q++) {
    B2_CHECK(hold_rank[q] >= 0 && hold_rank[q] < c->nranks, "b2_halo_set_exchange: holder rank out of range");
    if (hold_rank[q] == c->rank) continue;
    B2_CHECK(hold_pos[q] >= 0 && hold_spos[q] >= 0, "b2_halo_set_exchange: negative message position");
    B2_CHECK(!c->peer_base || (hold_pos[q] + kPeerScal < c->peer_slot && hold_spos[q] + kPeerScal < c->peer_slot),
             "b2_halo_set_exchange: message position %d / %d beyond the inbox slot (%lld cells)", hold_pos[q], hold_spos[q], (long long)c->peer_slot);
  }
  h->n_hold = nh;
  B2_TRY(b2_malloc(c, &h->hold_ptr, (size_t)h->n_if + 1));
  B2_TRY(b2_malloc(c, &h->hold_rank, (size_t)nh));
  B2_TRY(b2_malloc(c, &h->hold_pos, (size_t)nh));
  B2_TRY(b2_malloc(c, &h->hold_spos, (size_t)nh));
  B2_TRY(b2_upload(c, h->hold_ptr, hold_ptr, (size_t)h->n_if + 1));
  if (nh) {
    B2_TRY(b2_upload(c, h->hold_rank, hold_rank, (size_t)nh));
    B2_TRY(b2_upload(c, h->hold_pos, hold_pos, (size_t)nh));
    B2_TRY(b2_upload(c, h->hold_spos, hold_spos, (size_t)nh));
  }
  return 0;
}

/* reductions of v run over the owned entries only once a layout is attached (NULL detaches) */
int b2_vec_set_halo(b2_vec* v, const b2_halo* h) {
  B2_CHECK(!h || v->n >= h->n_local, "b2_vec_set_halo: vector shorter than the layout");
  v->halo = h;
  return 0;
}

}  // extern "C"

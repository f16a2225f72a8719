// pgb_pack.cu — the packed binary layout of posterior samples and the sampler checkpoint.
// The reference keeps PG_sample only in memory (Julia/src/PGopt.jl:23-29) and continues a chain through the x_prim
// keyword (particle_Gibbs.jl:114,126-128); this is the adjacent on-disk / on-wire format: one 64-byte header, then
// fixed-size records [A | Q | x_T | u_m1] — the same record the NCCL gather of pgb_multi_* moves — then, optionally,
// the particle blocks [w_m1 | x_m1].  Little-endian FP64 / Int64 throughout, version 1.
#include "pgb_internal.h"

using namespace pgb;

namespace {

struct PackHeader {  // 64 bytes
  char magic[8];
  uint32_t version, flags;
  int32_t n_x, n_u, n_phi, n_y;
  int64_t N, K;
  uint64_t aux[2];   // checkpoint: {Philox seed, chain_base | sweep_index << 32}
};
static_assert(sizeof(PackHeader) == 64, "header layout");

const char MAGIC_S[8] = {'P', 'G', 'B', '2', '0', '0', 'S', '\0'};
const char MAGIC_C[8] = {'P', 'G', 'B', '2', '0', '0', 'C', '\0'};

int64_t rec_doubles(const pgb_config& c) { return (int64_t)c.n_x * c.n_phi + (int64_t)c.n_x * c.n_x + c.n_x + c.n_u; }

}  // namespace

extern "C" int64_t pgb_samples_pack_bytes(const pgb_samples* s, int32_t with_particles) {
  if (!s) return 0;
  const pgb_config& c = s->hh.cfg;
  int64_t per = rec_doubles(c);
  if (with_particles) per += c.N + (int64_t)c.n_x * c.N;
  return (int64_t)sizeof(PackHeader) + s->K_active * per * 8;
}

extern "C" int pgb_samples_pack(pgb_samples* s, int32_t with_particles, void* buf, int64_t buf_bytes) {
  if (!s) return set_err(nullptr, PGB_ERR_INVALID, "sample set is NULL");
  pgb_handle* h = &s->hh;
  CK(h, cudaSetDevice(h->cfg.device));
  const pgb_config& c = h->cfg;
  const int64_t need = pgb_samples_pack_bytes(s, with_particles);
  if (!buf || buf_bytes < need) return set_err(h, PGB_ERR_INVALID, "buffer of %lld bytes needed", (long long)need);
  PackHeader hd;
  memset(&hd, 0, sizeof(hd));
  memcpy(hd.magic, MAGIC_S, 8);
  hd.version = 1;
  hd.flags = (with_particles ? 1u : 0u) | (s->have_xT ? 2u : 0u);  // bit 0: particle blocks follow, bit 1: x_T is valid
  hd.n_x = c.n_x; hd.n_u = c.n_u; hd.n_phi = c.n_phi; hd.n_y = c.n_y; hd.N = c.N; hd.K = s->K_active;
  memcpy(buf, &hd, sizeof(hd));
  double* out = reinterpret_cast<double*>(static_cast<char*>(buf) + sizeof(hd));
  const int64_t K = s->K_active, R = rec_doubles(c);
  const int nA = c.n_x * c.n_phi, nQ = c.n_x * c.n_x, nx = c.n_x, nu = c.n_u;
  cudaStream_t st = h->stream;
  // strided device-to-host copies: one 2-D copy per field straight into the record layout
  CK(h, cudaMemcpy2DAsync(out, (size_t)R * 8, s->A, (size_t)nA * 8, (size_t)nA * 8, (size_t)K, cudaMemcpyDeviceToHost, st));
  CK(h, cudaMemcpy2DAsync(out + nA, (size_t)R * 8, s->Q, (size_t)nQ * 8, (size_t)nQ * 8, (size_t)K, cudaMemcpyDeviceToHost, st));
  if (s->have_xT)
    CK(h, cudaMemcpy2DAsync(out + nA + nQ, (size_t)R * 8, s->xT, (size_t)nx * 8, (size_t)nx * 8, (size_t)K, cudaMemcpyDeviceToHost, st));
  CK(h, cudaMemcpy2DAsync(out + nA + nQ + nx, (size_t)R * 8, s->um, (size_t)nu * 8, (size_t)nu * 8, (size_t)K, cudaMemcpyDeviceToHost, st));
  if (with_particles) {
    double* pb = out + K * R;
    const int64_t P = c.N + (int64_t)nx * c.N;
    CK(h, cudaMemcpy2DAsync(pb, (size_t)P * 8, s->w, (size_t)c.N * 8, (size_t)c.N * 8, (size_t)K, cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpy2DAsync(pb + c.N, (size_t)P * 8, s->x, (size_t)nx * c.N * 8, (size_t)nx * c.N * 8, (size_t)K, cudaMemcpyDeviceToHost, st));
  }
  CK(h, cudaStreamSynchronize(st));
  if (!s->have_xT)
    for (int64_t k = 0; k < K; ++k)
      for (int d = 0; d < nx; ++d) out[k * R + nA + nQ + d] = 0.0;
  return PGB_OK;
}

extern "C" int pgb_samples_unpack(pgb_samples* s, const void* buf, int64_t buf_bytes) {
  if (!s) return set_err(nullptr, PGB_ERR_INVALID, "sample set is NULL");
  pgb_handle* h = &s->hh;
  CK(h, cudaSetDevice(h->cfg.device));
  const pgb_config& c = h->cfg;
  if (!buf || buf_bytes < (int64_t)sizeof(PackHeader)) return set_err(h, PGB_ERR_INVALID, "buffer too small for a header");
  PackHeader hd;
  memcpy(&hd, buf, sizeof(hd));
  if (memcmp(hd.magic, MAGIC_S, 8) != 0) return set_err(h, PGB_ERR_INVALID, "not a packed sample buffer (bad magic)");
  if (hd.version != 1) return set_err(h, PGB_ERR_INVALID, "unsupported packed-sample version %u", hd.version);
  if (hd.n_x != c.n_x || hd.n_u != c.n_u || hd.n_phi != c.n_phi)
    return set_err(h, PGB_ERR_INVALID, "packed samples have n_x=%d n_u=%d n_phi=%d, the set has %d %d %d", hd.n_x, hd.n_u, hd.n_phi,
                   c.n_x, c.n_u, c.n_phi);
  const bool parts = (hd.flags & 1u) != 0;
  if (parts && hd.N != c.N) return set_err(h, PGB_ERR_INVALID, "packed particle blocks have N=%lld, the set has N=%lld", (long long)hd.N, (long long)c.N);
  if (hd.K < 1 || hd.K > s->K) return set_err(h, PGB_ERR_INVALID, "packed K=%lld does not fit the set's capacity %lld", (long long)hd.K, (long long)s->K);
  const int64_t K = hd.K, R = rec_doubles(c);
  const int64_t P = c.N + (int64_t)c.n_x * c.N;
  const int64_t need = (int64_t)sizeof(PackHeader) + K * (R + (parts ? P : 0)) * 8;
  if (buf_bytes < need) return set_err(h, PGB_ERR_INVALID, "truncated buffer: %lld bytes needed", (long long)need);
  const double* in = reinterpret_cast<const double*>(static_cast<const char*>(buf) + sizeof(hd));
  const int nA = c.n_x * c.n_phi, nQ = c.n_x * c.n_x, nx = c.n_x, nu = c.n_u;
  cudaStream_t st = h->stream;
  CK(h, cudaMemcpy2DAsync(s->A, (size_t)nA * 8, in, (size_t)R * 8, (size_t)nA * 8, (size_t)K, cudaMemcpyHostToDevice, st));
  CK(h, cudaMemcpy2DAsync(s->Q, (size_t)nQ * 8, in + nA, (size_t)R * 8, (size_t)nQ * 8, (size_t)K, cudaMemcpyHostToDevice, st));
  CK(h, cudaMemcpy2DAsync(s->xT, (size_t)nx * 8, in + nA + nQ, (size_t)R * 8, (size_t)nx * 8, (size_t)K, cudaMemcpyHostToDevice, st));
  CK(h, cudaMemcpy2DAsync(s->um, (size_t)nu * 8, in + nA + nQ + nx, (size_t)R * 8, (size_t)nu * 8, (size_t)K, cudaMemcpyHostToDevice, st));
  if (parts) {
    const double* pb = in + K * R;
    CK(h, cudaMemcpy2DAsync(s->w, (size_t)c.N * 8, pb, (size_t)P * 8, (size_t)c.N * 8, (size_t)K, cudaMemcpyHostToDevice, st));
    CK(h, cudaMemcpy2DAsync(s->x, (size_t)nx * c.N * 8, pb + c.N, (size_t)P * 8, (size_t)nx * c.N * 8, (size_t)K, cudaMemcpyHostToDevice, st));
  }
  CK(h, cudaStreamSynchronize(st));
  s->K_active = K;
  // records without particle blocks can only be rolled out through their x_T (the star draw needs w_m1 / x_m1); with
  // particle blocks the star is re-drawn (same stream, same value) so that a stale x_T can never be used
  s->have_particles = parts;
  s->have_xT = !parts && (hd.flags & 2u) != 0;
  s->have_scen = false;
  return PGB_OK;
}

// ---------------------------------------------------------------------------------------------- sampler checkpoint
extern "C" int64_t pgb_checkpoint_bytes(const pgb_handle* h) {
  if (!h) return 0;
  const pgb_config& c = h->cfg;
  const int64_t per = (int64_t)c.n_x * c.n_phi + (int64_t)c.n_x * c.n_x + (int64_t)c.n_x * c.T;
  return (int64_t)sizeof(PackHeader) + (int64_t)c.n_chains * per * 8;
}

extern "C" int pgb_checkpoint_save(pgb_handle* h, void* buf, int64_t buf_bytes) {
  NEED(h);
  const pgb_config& c = h->cfg;
  const int64_t need = pgb_checkpoint_bytes(h);
  if (!buf || buf_bytes < need) return set_err(h, PGB_ERR_INVALID, "buffer of %lld bytes needed", (long long)need);
  for (char v : h->have_state)
    if (!v) return set_err(h, PGB_ERR_STATE, "a chain has no state yet");
  if (!h->have_rng || !h->fd.philox) return set_err(h, PGB_ERR_STATE, "a checkpoint needs Philox streams (injected streams cover one sweep)");
  PackHeader hd;
  memset(&hd, 0, sizeof(hd));
  memcpy(hd.magic, MAGIC_C, 8);
  hd.version = 1;
  hd.n_x = c.n_x; hd.n_u = c.n_u; hd.n_phi = c.n_phi; hd.n_y = c.n_y; hd.N = c.N; hd.K = c.n_chains;
  hd.flags = (uint32_t)c.T;
  hd.aux[0] = h->fd.seed;
  hd.aux[1] = (uint64_t)h->fd.chain_base | ((uint64_t)h->sweep_index << 32);
  memcpy(buf, &hd, sizeof(hd));
  double* out = reinterpret_cast<double*>(static_cast<char*>(buf) + sizeof(hd));
  const size_t nA = (size_t)c.n_x * c.n_phi, nQ = (size_t)c.n_x * c.n_x, nX = (size_t)c.n_x * c.T;
  const size_t per = nA + nQ + nX;
  CK(h, cudaStreamSynchronize(h->stream));
  for (int ch = 0; ch < c.n_chains; ++ch) {
    CK(h, cudaMemcpy(out + ch * per, h->fd.A + ch * nA, nA * 8, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(out + ch * per + nA, h->d_Q + ch * nQ, nQ * 8, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(out + ch * per + nA + nQ, h->fd.xprim + ch * nX, nX * 8, cudaMemcpyDeviceToHost));
  }
  return PGB_OK;
}

extern "C" int pgb_checkpoint_load(pgb_handle* h, const void* buf, int64_t buf_bytes) {
  NEED(h);
  const pgb_config& c = h->cfg;
  if (!buf || buf_bytes < (int64_t)sizeof(PackHeader)) return set_err(h, PGB_ERR_INVALID, "buffer too small for a header");
  PackHeader hd;
  memcpy(&hd, buf, sizeof(hd));
  if (memcmp(hd.magic, MAGIC_C, 8) != 0) return set_err(h, PGB_ERR_INVALID, "not a sampler checkpoint (bad magic)");
  if (hd.version != 1) return set_err(h, PGB_ERR_INVALID, "unsupported checkpoint version %u", hd.version);
  if (hd.n_x != c.n_x || hd.n_u != c.n_u || hd.n_phi != c.n_phi || hd.K != c.n_chains || hd.flags != (uint32_t)c.T)
    return set_err(h, PGB_ERR_INVALID, "checkpoint shape (n_x=%d n_u=%d n_phi=%d chains=%lld T=%u) does not match the handle", hd.n_x,
                   hd.n_u, hd.n_phi, (long long)hd.K, hd.flags);
  if (buf_bytes < pgb_checkpoint_bytes(h)) return set_err(h, PGB_ERR_INVALID, "truncated checkpoint");
  const double* in = reinterpret_cast<const double*>(static_cast<const char*>(buf) + sizeof(hd));
  const size_t nA = (size_t)c.n_x * c.n_phi, nQ = (size_t)c.n_x * c.n_x, nX = (size_t)c.n_x * c.T;
  const size_t per = nA + nQ + nX;
  const int64_t sweep_index = (int64_t)(hd.aux[1] >> 32);
  for (int ch = 0; ch < c.n_chains; ++ch) {
    int rc = pgb_set_state(h, ch, in + ch * per, in + ch * per + nA, in + ch * per + nA + nQ, sweep_index);
    if (rc) return rc;
  }
  return pgb_set_rng_philox(h, hd.aux[0], (uint32_t)(hd.aux[1] & 0xffffffffu));
}

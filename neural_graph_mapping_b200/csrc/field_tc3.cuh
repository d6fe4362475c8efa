// Three-slot form of the tcgen05 field kernel (included into field_tc.cu's translation unit).
//
// Why: with two tile slots, each owning its epilogue threads, a hidden layer costs a slot
// MMA 512 + synchronisation ~550 + epilogue ~700 cycles, so the tensor pipe sees 2 x 512 of work per
// ~1800 cycles (profiles/README.md).  A third tile in flight fills the gap, but TMEM (512 columns) only holds
// 2 x (128 fp32 accumulator + 64 fp16 A-operand columns) + 128: the third slot keeps its accumulator in TMEM and
// its A operand in SHARED memory (SS-mode MMA), and the register file only feeds two 256-thread epilogue teams.
// So the schedule is static and round-robin instead of slot-owned:
//
//   MMA order      (round r, layer l, slot s):  s fastest, then l, then r -- one issuer warp
//   epilogue jobs  in the same order, job j handled by team j mod 2 (two teams of 8 warps, each warp = one TMEM
//                  lane quadrant x one half of the columns); the epilogue itself is stateless (accumulator ->
//                  bias + ReLU -> fp16 A operand of the next layer, to TMEM for slots 0/1, to the swizzled
//                  shared-memory tile for slot 2)
//   front end      3 warps encode the NEXT round's rows into a per-slot staging buffer; the team that drains a
//                  slot's last layer installs the staged layer-0 operand for the slot's next tile
//
// Every hidden layer then offers the pipe 3 x 512 cycles of work per ~1700-cycle dependency chain.
// Field stage (MODE 1 of tc_kernel: points in, raw MLP outputs out), NeRF encoding, W = 128-class MLPs.
//
// STATUS: EXPERIMENT, selected only by NGM_TC3=1 (tests/test_gpu_tc.py keeps it correct).  Measured on the bench
// workload (19.66 M points, 4 x 128): 2.84 ms with three tiles in flight, 2.91 ms with two, 4.66 ms with one,
// against 2.16 ms of the production two-slot kernel whose epilogue threads are OWNED by their slot.  Removing the
// encoding work changes nothing, so neither the front end nor the tensor pipe (46% busy) is the limit: every slot's
// dependency chain stretches as slots are added (1.76 k -> 2.3 k -> 3.3 k cycles per layer) because the two shared
// teams serialise jobs that complete together and the SS-slot epilogue (x16 loads + shared-memory stores) is the
// slowest link.  Lessons kept in the code: (1) an mbarrier parity wait cannot tell phase k from k + 2 -- a waiter
// must see every phase of a barrier, hence one accumulator barrier per OWNING team; (2) one issuer warp per slot:
// a single issuer adds its ~600-cycle wake-up-to-commit latency to every MMA.  NGM_TC3_SLOTS=1|2|3 sets the
// number of tiles in flight.

constexpr int kThreads3 = 704;       // warps 0-2 MMA issuers of slots 0-2, warps 3-5 front end, warps 6-13 / 14-21 epilogue teams
constexpr int kFeThreads = 96;
constexpr int kTeamWarp0 = 6;
constexpr int kStageRowBytes = 128;  // staging row: up to 64 halves (EP <= 64)
constexpr int kAcBytes = 2 * 128 * 128;  // slot 2's A operand: 2 atoms (K = 128) x 128 rows x 128 B

struct Smem3 {
  uint64_t a0_ready[3];     // layer-0 A operand installed + accumulator drained (256 team threads)
  uint64_t a_ready[3];      // hidden-layer A operand stored (256 team threads)
  uint64_t d_ready[2][3];   // accumulator complete (tcgen05.commit), one barrier per OWNING team and slot: a parity
                            // wait cannot tell phase k from phase k + 2, so nobody may skip phases of a barrier --
                            // with per-team barriers every waiter sees every phase of the barriers it waits on
  uint64_t stage_full[3];   // staged rows written (96 front-end threads)
  uint64_t stage_empty[3];  // staged rows consumed (256 team threads)
  uint64_t w_ready;
  uint32_t tmem_base;
  uint32_t pad_;
};

// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// NeRF features of one point as EP/2 packed half2 words (same arithmetic as encode_nerf_to_tmem)
template <int OCT>
__device__ __forceinline__ void encode_nerf_words(float3 x, int start_octave, uint32_t (&w)[(6 * OCT + 15) / 16 * 8]) {
  constexpr int E = 6 * OCT;
  constexpr int EP = (E + 15) / 16 * 16;
  const float base = exp2f((float)start_octave);
  const float xs[3] = {x.x, x.y, x.z};
  float fe[EP];
#pragma unroll
  for (int i = E; i < EP; ++i) fe[i] = 0.0f;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float t0 = xs[d] * base;
    float s = 0.f, c = 1.f;
#pragma unroll
    for (int o = 0; o < OCT; ++o) {
      if (o % 4 == 0) {
        sincospi_fast(t0 * (float)(1 << o), s, c);
      } else {
        const float s2 = 2.0f * s * c;
        c = fmaf(-2.0f * s, s, 1.0f);
        s = s2;
      }
      fe[d * OCT + o] = s;
      fe[3 * OCT + d * OCT + o] = c;
    }
  }
#pragma unroll
  for (int j = 0; j < EP / 2; ++j) w[j] = ptx::pack_half2(fe[2 * j], fe[2 * j + 1]);
}

template <int OCT>
__global__ void __launch_bounds__(kThreads3, 1) tc3_kernel(const TcParams p) {
  constexpr int EPW = (6 * OCT + 15) / 16 * 8;  // layer-0 A operand words per row
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* wsm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);        // weight image (1024-B aligned)
  uint8_t* acs = wsm + (p.im.total_bytes + 1023) / 1024 * 1024;                // slot 2's A operand (1024-B aligned)
  uint8_t* stg = acs + kAcBytes;                                               // staging: [3][128][128 B]
  Smem3& sm = *reinterpret_cast<Smem3*>(stg + 3 * 128 * kStageRowBytes);
  const uint32_t wsm_addr = ptx::smem_u32(wsm), acs_addr = ptx::smem_u32(acs);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (warp == 0) ptx::tmem_alloc(&sm.tmem_base, kTmemCols);
  if (tid == 96) {
    for (int s = 0; s < 3; ++s) {
      ptx::mbar_init(&sm.a0_ready[s], 256);
      ptx::mbar_init(&sm.a_ready[s], 256);
      ptx::mbar_init(&sm.d_ready[0][s], 1);
      ptx::mbar_init(&sm.d_ready[1][s], 1);
      ptx::mbar_init(&sm.stage_full[s], kFeThreads);
      ptx::mbar_init(&sm.stage_empty[s], 256);
    }
    ptx::mbar_init(&sm.w_ready, 1);
    ptx::fence_mbar_init();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  const long long t_begin = p.total_tiles * blockIdx.x / gridDim.x;
  const long long t_end = p.total_tiles * (blockIdx.x + 1) / gridDim.x;
  const int L = p.L, W = p.W;
  // TMEM columns: slot 0: D [0,128) A [128,192); slot 1: D [192,320) A [320,384); slot 2: D [384,512)
  auto d_col = [](int s) { return s == 0 ? 0u : (s == 1 ? 192u : 384u); };

  uint32_t w_phase = 0;
  // completed-phase counters of the per-slot barriers (every role tracks all of them: jobs of the other team
  // flip phases too)
  int n_a0[3] = {0, 0, 0}, n_a[3] = {0, 0, 0}, n_d[3] = {0, 0, 0}, n_full[3] = {0, 0, 0}, n_empty[3] = {0, 0, 0};
  int job = 0;  // global epilogue-job counter (team = job & 1)

  long long t = t_begin;
  while (t < t_end) {
    const long long f = t / p.tiles_per_field;
    long long seg_end = (f + 1) * p.tiles_per_field;
    if (seg_end > t_end) seg_end = t_end;
    const int ntiles = (int)(seg_end - t);
    const long long tile0_in_field = t - f * p.tiles_per_field;
    const int NS = p.knn_k;  // tiles in flight (3; 1 or 2 for experiments)
    const int rounds = (ntiles + NS - 1) / NS;
    const long long slot = p.field_slots ? p.field_slots[f] : f;

    if (tid == 0) {  // stage this field's weight image (TMA engine)
      const uint8_t* src = p.images + (size_t)f * p.im.total_bytes;
      ptx::mbar_arrive_expect_tx(&sm.w_ready, p.im.total_bytes);
      for (uint32_t o = 0; o < p.im.total_bytes; o += 32768) {
        const uint32_t n = p.im.total_bytes - o < 32768 ? p.im.total_bytes - o : 32768;
        ptx::bulk_g2s(wsm + o, src + o, n, &sm.w_ready);
      }
    }

    if (warp < 3) {
      // ===================== MMA issuer of slot `warp` (its tiles: warp, warp + NS, ...) =====================
      // One issuer per slot: the ~600 cycles between a barrier wake-up and the commit (fence, descriptor set-up,
      // MMA queue back-pressure) of the three slots overlap instead of adding up in one thread.
      const int s = warp;
      ptx::mbar_wait(&sm.w_ready, w_phase);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t wsm_u = __shfl_sync(0xffffffffu, wsm_addr, 0);
      const uint32_t acs_u = __shfl_sync(0xffffffffu, acs_addr, 0);
      const uint32_t d_addr = tmem_u + d_col(s);
      int na0 = n_a0[0], na = n_a[0];  // this issuer only tracks its own slot (kept in element 0)
      const int boot = ntiles < NS ? ntiles : NS;  // install-only jobs that open the segment
      if (s < NS) {
        for (int r = 0; NS * r + s < ntiles; ++r) {
          const int active = ntiles - NS * r < NS ? ntiles - NS * r : NS;  // tiles of this round
          for (int l = 0; l <= L; ++l) {
            // global epilogue-job index of (r, l, s) -> the team that owns it (job parity)
            const int owner = (job + boot + r * NS * (L + 1) + l * active + s) & 1;
            const TcLayer y = p.im.layer[l];
            const uint64_t bdesc0 = ptx::make_smem_desc_sw128(wsm_u + y.off);
            const uint32_t idesc = ptx::make_idesc_f16(y.n_pad);
            const int ksteps = y.k_pad / 16;
            if (l == 0) { ptx::mbar_wait_lean(&sm.a0_ready[s], na0 & 1); ++na0; }
            else        { ptx::mbar_wait_lean(&sm.a_ready[s], na & 1);   ++na; }
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              if (s < 2) {
                issue_layer(d_addr, d_addr + kACol, bdesc0, (uint32_t)y.n_pad * 8u, idesc, ksteps);
              } else {
                const uint64_t adesc0 = ptx::make_smem_desc_sw128(acs_u);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                  if (ks < ksteps) {
                    const uint64_t bd = bdesc0 + (uint64_t)((ks & 3) * 2u + (ks >> 2) * ((uint32_t)y.n_pad * 8u));
                    const uint64_t ad = adesc0 + (uint64_t)((ks & 3) * 2u + (ks >> 2) * 1024u);
                    mma_f16_ss(d_addr, ad, bd, idesc, ks > 0 ? 1u : 0u);
                  }
                }
              }
              ptx::mma_commit(&sm.d_ready[owner][s]);
            }
            __syncwarp();
          }
        }
      }
      n_a0[0] = na0;
      n_a[0] = na;
      job += boot + ntiles * (L + 1);
    } else if (warp < kTeamWarp0) {
      // ===================== front end: rows of the next tiles -> staging =====================
      const int ft = tid - 96;  // 0..95: rows ft (and ft + 96 for the first 32 threads)
      for (int ti = 0; ti < ntiles; ++ti) {
        const int s = ti % NS;
        // the staging buffer of slot s is free once the previous tile's rows were installed
        if (s == 0) { ptx::mbar_wait(&sm.stage_empty[0], (n_empty[0] + 1) & 1); ++n_empty[0]; }
        else if (s == 1) { ptx::mbar_wait(&sm.stage_empty[1], (n_empty[1] + 1) & 1); ++n_empty[1]; }
        else { ptx::mbar_wait(&sm.stage_empty[2], (n_empty[2] + 1) & 1); ++n_empty[2]; }
        for (int row = ft; row < 128; row += kFeThreads) {
          const long long gp = (tile0_in_field + ti) * 128 + row;
          float3 fx = make_float3(0.f, 0.f, 0.f);
          if (gp < p.points_per_field) {
            const float* src = p.points + (f * p.points_per_field + gp) * 3;
            float3 x = make_float3(__ldg(src), __ldg(src + 1), __ldg(src + 2));
            if (p.positions) {
              const float* c = p.positions + slot * 3;
              const float* q = p.orientations + slot * 4;
              x = make_float3(x.x - __ldg(c), x.y - __ldg(c + 1), x.z - __ldg(c + 2));
              x = quat_inv_rotate(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), x);
            }
            fx = scale_local(x, p.scale_mode, p.field_radius);
          }
          uint32_t w[EPW];
          encode_nerf_words<OCT>(fx, p.nerf_start, w);
          uint4* dst = reinterpret_cast<uint4*>(stg + (s * 128 + row) * kStageRowBytes);
#pragma unroll
          for (int j = 0; j < EPW / 4; ++j) dst[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        }
        ptx::mbar_arrive(&sm.stage_full[s]);
      }
    } else {
      // ===================== epilogue teams =====================
      const int team = (warp - kTeamWarp0) >> 3;
      const int h = ((warp - kTeamWarp0) >> 2) & 1;  // column half
      const int qwarp = warp & 3;           // TMEM lane quadrant
      const int row = qwarp * 32 + lane;
      const uint32_t lane_base = tmem_base + ((uint32_t)(qwarp * 32) << 16);
      const uint32_t* bias2 = reinterpret_cast<const uint32_t*>(wsm + p.im.bias_h2_off);
      const float* bias_last = reinterpret_cast<const float*>(wsm + p.im.bias_last_off);
      const int w0 = ((W / 16 + 1) / 2) * 16;
      const int my_c0 = h ? w0 : 0, my_n = h ? W - w0 : w0;
      ptx::mbar_wait(&sm.w_ready, w_phase);

      // install the staged layer-0 A operand of slot s (this thread: its row, its half of the words)
      auto install = [&](int s) {  // (the caller has waited for stage_full[s])
        const uint32_t* src = reinterpret_cast<const uint32_t*>(stg + (s * 128 + row) * kStageRowBytes);
        constexpr int HW = EPW / 2;  // words per half (EPW is a multiple of 8)
        uint32_t w[HW];
#pragma unroll
        for (int j = 0; j < HW; j += 4) {
          const uint4 v = *reinterpret_cast<const uint4*>(src + h * HW + j);
          w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
        }
        if (s < 2) {
          ptx::tmem_store_n<HW>(lane_base + d_col(s) + kACol + h * HW, w);
          ptx::tc_wait_st();
          ptx::tc_fence_before();
        } else {  // K-major SWIZZLE_128B rows: 16-B chunk c of row r lives at chunk c ^ (r & 7)
#pragma unroll
          for (int j = 0; j < HW; j += 4) {
            const int chunk = (h * HW + j) >> 2;
            *reinterpret_cast<uint4*>(acs + row * 128 + ((chunk ^ (row & 7)) << 4)) = make_uint4(w[j], w[j + 1], w[j + 2], w[j + 3]);
          }
          ptx::fence_proxy_async();
        }
        ptx::mbar_arrive(&sm.stage_empty[s]);
        ptx::mbar_arrive(&sm.a0_ready[s]);
      };

      // bootstrap: the first round's layer-0 operands (three install-only jobs)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        if (s < NS && s < ntiles) {
          if ((job & 1) == team) {
            ptx::mbar_wait_lean(&sm.stage_full[s], n_full[s] & 1);
            install(s);
          }
          ++n_full[s];
          ++job;
        }
      }
      for (int r = 0; r < rounds; ++r) {
        for (int l = 0; l <= L; ++l) {
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const int ti = NS * r + s;
            if (s >= NS || ti >= ntiles) continue;
            const bool mine = (job & 1) == team;
            ++job;
            const bool next_tile = l == L && ti + NS < ntiles;
            if (next_tile) ++n_full[s];
            if (!mine) continue;
            ptx::mbar_wait_lean(&sm.d_ready[team][s], n_d[s] & 1);  // n_d: OWN jobs on this slot so far
            ++n_d[s];
            ptx::tc_fence_after();
            const uint32_t d_addr = lane_base + d_col(s);
            if (l < L) {
              // ---------- hidden layer: accumulator -> bias + ReLU -> next A operand ----------
              const uint32_t* b2 = bias2 + l * (W / 2);
              if (s < 2) {
                if (my_n > 0) hidden_epilogue(d_addr, d_addr + kACol, b2, my_c0, my_n);
                ptx::tc_wait_st();
                ptx::tc_fence_before();
              } else {
                for (int c = my_c0; c < my_c0 + my_n; c += 16) {
                  uint32_t v[16];
                  ptx::tmem_ld16(d_addr + c, v);
                  ptx::tc_wait_ld();
                  uint32_t w[8];
                  cvt16(v, b2 + c / 2, w);
                  // columns c..c+15 = K elements of atom c/64, 16-B chunks (c%64)/8 and +1
                  uint8_t* base = acs + (c >> 6) * (128 * 128) + row * 128;
                  const int ck = (c & 63) >> 3;
                  *reinterpret_cast<uint4*>(base + ((ck ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                  *reinterpret_cast<uint4*>(base + (((ck + 1) ^ (row & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
                }
                ptx::tc_fence_before();
                ptx::fence_proxy_async();
              }
              ptx::mbar_arrive(&sm.a_ready[s]);
            } else {
              // ---------- last layer: outputs to HBM, then the slot's next layer-0 operand ----------
              if (h == 0) {
                const long long gp = (tile0_in_field + ti) * 128 + row;
                const bool valid = gp < p.points_per_field;
                float* o = p.out + (f * p.points_per_field + gp) * p.dim_out;
                for (int c = 0; c < p.im.layer[L].n_pad; c += 16) {
                  uint32_t v[16];
                  ptx::tmem_ld16(d_addr + c, v);
                  ptx::tc_wait_ld();
                  if (valid) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                      if (c + i < p.dim_out) o[c + i] = __uint_as_float(v[i]) + bias_last[c + i];
                  }
                }
                ptx::tc_fence_before();
              }
              if (next_tile) {
                // staged rows of the slot's next tile: phase n_full - 1 (counted above).  Skipping the phases
                // installed by the other team is safe HERE: the single staging buffer hand-shakes through
                // stage_empty, so this barrier is never more than one phase ahead of or behind its waiter.
                ptx::mbar_wait_lean(&sm.stage_full[s], (n_full[s] - 1) & 1);
                install(s);
              }
            }
          }
        }
      }
    }
    w_phase ^= 1;
    t = seg_end;
    ptx::fence_proxy_async();  // generic-proxy reads of the image before the next bulk copy overwrites it
    __syncthreads();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

size_t tc3_smem_bytes(const TcImage& im) {
  return 1024 + (im.total_bytes + 1023) / 1024 * 1024 + kAcBytes + 3 * 128 * kStageRowBytes + sizeof(Smem3) + 128;
}

// NGM_TC3=1 selects the three-slot kernel for the shapes it covers (field stage, NeRF encoding)
bool tc3_enabled() {
  const char* e = getenv("NGM_TC3");
  return e && e[0] == '1';
}

bool tc3_supported(const NgmFieldDesc& fd, const TcImage& im) {
  return fd.encoding == NGM_ENC_NERF && nerf_octaves_supported(fd.nerf_num_octaves) && fd.num_layers >= 1 &&
         tc3_smem_bytes(im) <= 227 * 1024;
}

int launch_tc3(const TcParams& p_in, int octaves, int grid, cudaStream_t stream) {
  TcParams p = p_in;
  const char* e = getenv("NGM_TC3_SLOTS");
  p.knn_k = e ? atoi(e) : 3;
  if (p.knn_k < 1 || p.knn_k > 3) p.knn_k = 3;
  const size_t smem = tc3_smem_bytes(p.im);
  auto go = [&](auto kernel) -> int {
    NGM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, kThreads3, smem, stream>>>(p);
    return check_launch("tc3_kernel");
  };
  return octaves == 4 ? go(tc3_kernel<4>) : go(tc3_kernel<8>);
}

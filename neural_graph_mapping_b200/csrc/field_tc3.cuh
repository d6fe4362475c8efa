// Three-slot form of the tcgen05 field kernel (included into field_tc.cu's translation unit).
//
// Why: with two tile slots a hidden layer costs a slot MMA 512 + synchronisation ~550 + epilogue ~700 cycles,
// so the tensor pipe sees 2 x 512 cycles of work per ~1800 (profiles/README.md).  A third tile in flight fills
// the gap.  TMEM (512 columns) holds 2 x (128 fp32 accumulator + 64 fp16 A-operand columns) + 128, so the third
// slot keeps its accumulator in TMEM and its A operand in SHARED memory (SS-mode tcgen05.mma), and there are no
// spare columns to stage the next tile's layer-0 operand: it is staged in shared memory and installed by the
// slot's own threads when the last layer has drained.
//
//   warps 0-2                   MMA issuer of slot 0 / 1 / 2 (blocking mbarrier wait, elected lane issues, commit)
//   warps 4-11 / 12-19 / 20-27  the 256 threads that OWN slot 0 / 1 / 2: two threads per row (TMEM lane), each
//                               half of the columns of every epilogue; the h = 1 half also encodes the slot's next
//                               tile into the staging buffer while the MMAs run
//
// 896 threads leave 72 registers per thread.
//
// STATUS: EXPERIMENT, selected only by NGM_TC3=1 (tests/test_gpu_tc.py keeps it correct).  Measured on the bench
// workload (19.66 M points, 4 x 128 MLP), field stage: 5.74 / 3.45 / 3.07 ms with 1 / 2 / 3 tiles in flight against
// 2.16 ms of the production two-slot kernel.  The extra tile overlaps as intended (1.9x from one to three slots),
// but every slot's dependency chain is longer here than in the two-slot kernel (2.2 k cycles per layer with one
// tile in flight against 1.8 k: the layer-0 operand goes through shared memory and is installed on the critical
// path, 28 warps share the four schedulers) and it stretches further as slots are added (3.5 k with three: the
// epilogues' TMEM loads and stores compete with the accumulator traffic of the other slots' MMAs, which now keep
// the tensor pipe busy most of the time).  A first variant with two epilogue teams SHARED by the three slots in a
// static round-robin order (4.66 / 2.91 / 2.84 ms) lost to the serialisation of jobs that complete together.
// Two lessons from it are kept: an mbarrier parity wait cannot tell phase k from phase k + 2, so every waiter must
// see every phase of the barriers it waits on; and one issuer per slot, because a single issuer adds its
// ~600-cycle wake-up-to-commit latency to every MMA.
// Field stage only (MODE 1 of tc_kernel: points in, raw MLP outputs out), NeRF encoding;
// NGM_TC3_SLOTS=1|2|3 sets the number of tiles in flight.

constexpr int kThreads3 = 896;
constexpr int kTeamWarp0 = 4;
constexpr int kStageRowBytes = 128;      // staging row: up to 64 halves (EP <= 64)
constexpr int kAcBytes = 2 * 128 * 128;  // slot 2's A operand: 2 atoms (K = 128) x 128 rows x 128 B

struct Smem3 {
  uint64_t a0_ready[3];    // layer-0 A operand installed + accumulator drained (256 threads of the slot)
  uint64_t a_ready[3];     // hidden-layer A operand stored (256)
  uint64_t d_ready[3];     // accumulator complete (tcgen05.commit)
  uint64_t stage_full[3];  // staged rows of the next tile written (the 128 h = 1 threads)
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

// NeRF features of one point -> one staging row, one input dimension at a time (OCT sines + OCT cosines live in
// registers instead of all 6 * OCT features).  Row layout = the A operand: [sin: dim-major | cos: dim-major | 0 pad].
template <int OCT>
__device__ __forceinline__ void encode_nerf_row(float3 x, int start_octave, uint32_t* row_words) {
  constexpr int E = 6 * OCT, EP = (E + 15) / 16 * 16;
  const float base = exp2f((float)start_octave);
  const float xs[3] = {x.x, x.y, x.z};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float t0 = xs[d] * base;
    float sv[OCT], cv[OCT];
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
      sv[o] = s;
      cv[o] = c;
    }
#pragma unroll
    for (int j = 0; j < OCT / 2; ++j) {
      row_words[d * (OCT / 2) + j] = ptx::pack_half2(sv[2 * j], sv[2 * j + 1]);
      row_words[3 * (OCT / 2) + d * (OCT / 2) + j] = ptx::pack_half2(cv[2 * j], cv[2 * j + 1]);
    }
  }
#pragma unroll
  for (int j = E / 2; j < EP / 2; ++j) row_words[j] = 0u;
}

template <int OCT>
__global__ void __launch_bounds__(kThreads3, 1) tc3_kernel(const TcParams p) {
  constexpr int EPW = (6 * OCT + 15) / 16 * 8;  // layer-0 A operand words per row
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* wsm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);  // weight image (1024-B aligned)
  uint8_t* acs = wsm + (p.im.total_bytes + 1023) / 1024 * 1024;          // slot 2's A operand (1024-B aligned)
  uint8_t* stg = acs + kAcBytes;                                         // staging: [3][128][128 B]
  Smem3& sm = *reinterpret_cast<Smem3*>(stg + 3 * 128 * kStageRowBytes);
  const uint32_t wsm_addr = ptx::smem_u32(wsm), acs_addr = ptx::smem_u32(acs);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (warp == 0) ptx::tmem_alloc(&sm.tmem_base, kTmemCols);
  if (tid == 96) {
    for (int s = 0; s < 3; ++s) {
      ptx::mbar_init(&sm.a0_ready[s], 256);
      ptx::mbar_init(&sm.a_ready[s], 256);
      ptx::mbar_init(&sm.d_ready[s], 1);
      ptx::mbar_init(&sm.stage_full[s], 128);
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
  const int NS = p.knn_k;  // tiles in flight (3; 1 or 2 for experiments)
  // TMEM columns: slot 0: D [0,128) A [128,192); slot 1: D [192,320) A [320,384); slot 2: D [384,512)
  auto d_col = [](int s) { return s == 0 ? 0u : (s == 1 ? 192u : 384u); };

  uint32_t w_phase = 0;
  int n_a0 = 0, n_a = 0, n_d = 0, n_full = 0;  // completed phases of this thread's slot barriers

  long long t = t_begin;
  while (t < t_end) {
    const long long f = t / p.tiles_per_field;
    long long seg_end = (f + 1) * p.tiles_per_field;
    if (seg_end > t_end) seg_end = t_end;
    const int ntiles = (int)(seg_end - t);
    const long long tile0_in_field = t - f * p.tiles_per_field;
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
      const int s = warp;
      ptx::mbar_wait(&sm.w_ready, w_phase);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t wsm_u = __shfl_sync(0xffffffffu, wsm_addr, 0);
      const uint32_t acs_u = __shfl_sync(0xffffffffu, acs_addr, 0);
      const uint32_t d_addr = tmem_u + d_col(s);
      if (s < NS) {
        for (int ti = s; ti < ntiles; ti += NS) {
          for (int l = 0; l <= L; ++l) {
            const TcLayer y = p.im.layer[l];
            const uint64_t bdesc0 = ptx::make_smem_desc_sw128(wsm_u + y.off);
            const uint32_t idesc = ptx::make_idesc_f16(y.n_pad);
            const int ksteps = y.k_pad / 16;
            if (l == 0) { ptx::mbar_wait_lean(&sm.a0_ready[s], n_a0 & 1); ++n_a0; }
            else        { ptx::mbar_wait_lean(&sm.a_ready[s], n_a & 1);   ++n_a; }
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
              ptx::mma_commit(&sm.d_ready[s]);
            }
            __syncwarp();
          }
        }
      }
    } else if (warp >= kTeamWarp0) {
      // ===================== the 256 threads of slot s =====================
      const int s = (warp - kTeamWarp0) >> 3;
      const int h = ((warp - kTeamWarp0) >> 2) & 1;  // column half; h = 1 also runs the front end
      const int qwarp = warp & 3;                    // TMEM lane quadrant
      const int row = qwarp * 32 + lane;
      const uint32_t d_addr = tmem_base + ((uint32_t)(qwarp * 32) << 16) + d_col(s);
      const uint32_t* bias2 = reinterpret_cast<const uint32_t*>(wsm + p.im.bias_h2_off);
      const float* bias_last = reinterpret_cast<const float*>(wsm + p.im.bias_last_off);
      const int w0 = ((W / 16 + 1) / 2) * 16;
      const int my_c0 = h ? w0 : 0, my_n = h ? W - w0 : w0;
      uint32_t* my_stage = reinterpret_cast<uint32_t*>(stg + (s * 128 + row) * kStageRowBytes);

      // front end of tile ti (h == 1 threads, one row each): point -> local -> NeRF features -> staging row
      auto fe = [&](int ti) {
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
        encode_nerf_row<OCT>(fx, p.nerf_start, my_stage);
        ptx::mbar_arrive(&sm.stage_full[s]);
      };
      // install the staged layer-0 A operand (this thread: its row, its half of the words), then signal the issuer
      auto install = [&]() {
        ptx::mbar_wait_lean(&sm.stage_full[s], n_full & 1);
        ++n_full;
        constexpr int HW = EPW / 2;  // words per half (a multiple of 4)
        uint32_t w[HW];
#pragma unroll
        for (int j = 0; j < HW; j += 4) {
          const uint4 v = *reinterpret_cast<const uint4*>(my_stage + h * HW + j);
          w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
        }
        if (s < 2) {
          ptx::tmem_store_n<HW>(d_addr + kACol + h * HW, w);
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
        ptx::mbar_arrive(&sm.a0_ready[s]);
      };

      if (s < NS && s < ntiles) {
        ptx::mbar_wait(&sm.w_ready, w_phase);  // biases live in the image
        if (h == 1) fe(s);
        install();
        for (int ti = s; ti < ntiles; ti += NS) {
          const bool has_next = ti + NS < ntiles;
          for (int l = 0; l <= L; ++l) {
            ptx::mbar_wait_lean(&sm.d_ready[s], n_d & 1);
            ++n_d;
            ptx::tc_fence_after();
            if (l < L) {
              // ---------- hidden layer: accumulator -> bias + ReLU -> next A operand (16 columns per step) ----------
              const uint32_t* b2 = bias2 + l * (W / 2);
              int c = my_c0;
              const int cend = my_c0 + my_n;
              for (; c + 32 <= cend; c += 32) {  // 32 columns per TMEM load: one exposed load latency per step
                uint32_t v[32];
                ptx::tmem_ld32(d_addr + c, v);
                ptx::tc_wait_ld();
                uint32_t w[16];
                cvt16(v, b2 + c / 2, w);
                cvt16(v + 16, b2 + c / 2 + 8, w + 8);
                if (s < 2) {
                  ptx::tmem_st16(d_addr + kACol + c / 2, w);
                } else {  // columns c..c+31 = K elements of atom c/64, four 16-B chunks from (c%64)/8
                  uint8_t* base = acs + (c >> 6) * (128 * 128) + row * 128;
                  const int ck = (c & 63) >> 3;
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(base + (((ck + j) ^ (row & 7)) << 4)) =
                        make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                }
              }
              for (; c < cend; c += 16) {
                uint32_t v[16];
                ptx::tmem_ld16(d_addr + c, v);
                ptx::tc_wait_ld();
                uint32_t w[8];
                cvt16(v, b2 + c / 2, w);
                if (s < 2) {
                  ptx::tmem_st8(d_addr + kACol + c / 2, w);
                } else {
                  uint8_t* base = acs + (c >> 6) * (128 * 128) + row * 128;
                  const int ck = (c & 63) >> 3;
                  *reinterpret_cast<uint4*>(base + ((ck ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                  *reinterpret_cast<uint4*>(base + (((ck + 1) ^ (row & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
                }
              }
              if (s < 2) ptx::tc_wait_st(); else ptx::fence_proxy_async();
              ptx::tc_fence_before();
              ptx::mbar_arrive(&sm.a_ready[s]);
              // front end of the slot's next tile, in the wait for this layer's successor MMA
              if (l == 0 && h == 1 && has_next) fe(ti + NS);
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
              if (has_next) install();
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

// Host side of the HG stage: weight intake (eval-mode BatchNorm folded into the convs), B-operand packing for
// gconv_kernel, per-resolution workspace + launch plan, FP32 parity path.  Included by engine.cu inside namespace hdrtv.
// Reference: Hallucination_arch.py:53-137 (network), :240-275 (the BatchNorm fold), HG_Composite_arch.py:77-107.
#pragma once

static const char* kHgBnBlocks[] = {"conv1", "conv2", "conv3_1", "conv3_2", "conv4_1", "conv4_2", "conv5_1", "conv5_2", "conv_code1", "conv_code2"};
struct HgLayerSpec { const char* name; int cin, cout, ks; };
static const HgLayerSpec kHgLayers[] = {
    {"conv1.0", 3, 64, 3},        {"conv2.0", 64, 128, 3},      {"conv3_1.0", 128, 256, 3},   {"conv3_2.0", 256, 256, 3},
    {"conv4_1.0", 256, 512, 3},   {"conv4_2.0", 512, 512, 3},   {"conv5_1.0", 512, 512, 3},   {"conv5_2.0", 512, 512, 3},
    {"conv_code1.0", 512, 512, 3}, {"conv_code2.0", 512, 512, 3}, {"Up_conv1.0", 512, 2048, 3}, {"Up_conv2.0", 512, 2048, 3},
    {"Up_conv3.0", 256, 1024, 3}, {"Up_conv4.0", 128, 512, 3},  {"Up_conv5.0", 64, 256, 3},   {"conv6", 1024, 512, 1},
    {"conv7", 1024, 256, 1},      {"conv8", 512, 128, 1},       {"conv9", 256, 64, 1},        {"conv10", 128, 3, 1},
    {"conv_last", 6, 3, 1}};

static void hg_release_ws(Ctx* c) {
  for (void* p : c->hg.ws) cudaFree(p);
  c->hg.ws.clear();
  c->hg.ws_bytes = 0;
  c->hg.plan.clear();
  c->hg.f32.clear();
  c->hg.t.clear();
  c->hg.H = c->hg.W = 0;
  c->hg.proc_out = nullptr;
  c->hg.proc_H = c->hg.proc_W = 0;
  c->hg.d_part = nullptr;
  c->hg.d_gate = nullptr;
}
static void hg_release_weights(Ctx* c) {
  for (void* p : c->hg.wallocs) cudaFree(p);
  c->hg.wallocs.clear();
  c->hg.wpk.clear();
  c->hg.wd.clear();
  c->hg.w.clear();
  c->hg.d_tail = nullptr;
  c->hg.d_dot_up = c->hg.d_dot_skip = nullptr;
  c->hg.has = false;
}
template <typename T>
static T* hg_ws_alloc(Ctx* c, size_t n, bool zero) {
  void* p = nullptr;
  if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return nullptr;
  if (zero) cudaMemset(p, 0, n * sizeof(T));
  c->hg.ws.push_back(p);
  c->hg.ws_bytes += n * sizeof(T);
  return static_cast<T*>(p);
}
template <typename T>
static T* hg_w_upload(Ctx* c, const T* host, size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return nullptr;
  cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice);
  c->hg.wallocs.push_back(p);
  return static_cast<T*>(p);
}

// P8 tensor with row slack: gconv tiles are 4 rows tall and read one halo row above / below, so the allocation covers
// P8 rows 0 .. rup(H, 4) + 1; rows past H stay zero (no epilogue writes them) = the convolution's zero padding.
static P8 hg_make_p8(Ctx* c, int C, int H, int Wd) {
  P8 t;
  t.chunks = (C + 7) / 8;
  t.H = H;
  t.W = Wd;
  t.parity = 0;
  t.Wp = rup(Wd, kTileM) + 8;
  const size_t rows = static_cast<size_t>(rup(H, kGRows)) + 2;
  t.base = hg_ws_alloc<__half>(c, rows * t.chunks * t.Wp * 8, true);
  return t;
}

// B operand of gconv_kernel: per N tile, per K group, g_steps blocks of [16 x NT] (bpack_index layout), then one bias block
static std::vector<__half> hg_pack(const HostTensor& w, const HostTensor& b, int kind, int NT) {
  const int Cout = static_cast<int>(w.shape[0]), Cin = static_cast<int>(w.shape[1]);
  const int taps = w.shape.size() == 4 ? static_cast<int>(w.shape[2] * w.shape[3]) : 1;
  const int gc = g_group_channels(kind), steps = g_steps(kind);
  const int ntiles = (Cout + NT - 1) / NT, kgroups = (Cin + gc - 1) / gc;
  const size_t blk = static_cast<size_t>(NT) * 16;                    // halves per block
  const size_t tile = (static_cast<size_t>(kgroups) * steps + 1) * blk;
  std::vector<__half> out(tile * ntiles, __float2half(0.f));
  const float* wd = w.v.data();
  auto wv = [&](int co, int ci, int tap) -> float {
    if (co >= Cout || ci >= Cin || tap >= taps) return 0.f;
    return wd[(static_cast<size_t>(co) * Cin + ci) * taps + tap];
  };
  for (int nt = 0; nt < ntiles; ++nt) {
    __half* tb = out.data() + tile * nt;
    for (int kg = 0; kg < kgroups; ++kg)
      for (int s = 0; s < steps; ++s) {
        __half* bb = tb + (static_cast<size_t>(kg) * steps + s) * blk;
        for (int h = 0; h < 2; ++h) {
          int tap = 0, cbase = 0;
          if (kind == G_3x3) { tap = s; cbase = kg * 16 + h * 8; }
          else if (kind == G_3x3_C8) {
            const int dy = s / 2;
            if (s % 2 == 0) tap = dy * 3 + h;
            else if (h == 0) tap = dy * 3 + 2;
            else continue;                                            // zero weights
            cbase = 0;
          } else { tap = 0; cbase = kg * 64 + (2 * s + h) * 8; }
          for (int n = 0; n < NT; ++n)
            for (int e = 0; e < 8; ++e)
              bb[bpack_index(NT, 0, n, h * 8 + e)] = __float2half(wv(nt * NT + n, cbase + e, tap));
        }
      }
    __half* bias = tb + static_cast<size_t>(kgroups) * steps * blk;
    for (int n = 0; n < NT; ++n) {
      const float bv = nt * NT + n < Cout ? b.v[nt * NT + n] : 0.f;
      const __half hi = __float2half(bv);
      bias[bpack_index(NT, 0, n, 0)] = hi;
      bias[bpack_index(NT, 0, n, 1)] = __float2half(bv - __half2float(hi));
    }
  }
  return out;
}

struct HgShape { int kind, NT; };
static HgShape hg_layer_shape(const HgLayerSpec& L) {
  if (L.ks == 3) return L.cin == 3 ? HgShape{G_3x3_C8, 64} : HgShape{G_3x3, 128};
  return HgShape{G_1x1, L.cout >= 128 ? 128 : 64};
}

static int hg_set_weights(Ctx* c, const hdrtv_tensor_desc* t, int n) {
  hg_release_ws(c);
  hg_release_weights(c);
  std::map<std::string, HostTensor> raw;
  for (int i = 0; i < n; ++i) {
    HostTensor ht;
    size_t cnt = 1;
    for (int d = 0; d < t[i].ndim; ++d) { ht.shape.push_back(t[i].shape[d]); cnt *= static_cast<size_t>(t[i].shape[d]); }
    ht.v.assign(t[i].data, t[i].data + cnt);
    std::string key = t[i].name;
    if (key.rfind("module.", 0) == 0) key = key.substr(7);
    if (key.rfind("hg.", 0) == 0) key = key.substr(3);
    raw[key] = std::move(ht);
  }
  // strict key / shape check (model.hg.load_state_dict(hg_state, strict=True), hdrtvnet_torch.py:2143)
  for (const HgLayerSpec& L : kHgLayers) {
    const std::string wk = std::string(L.name) + ".weight", bk = std::string(L.name) + ".bias";
    if (!raw.count(wk) || !raw.count(bk)) return fail(c, "hdrtv_set_hg_weights: missing key " + wk);
    const HostTensor& w = raw.at(wk);
    if (w.shape.size() != 4 || w.shape[0] != L.cout || w.shape[1] != L.cin || w.shape[2] != L.ks || w.shape[3] != L.ks)
      return fail(c, "hdrtv_set_hg_weights: unexpected shape for " + wk + " (the PixelShuffle / FusedBN Hallucination_Generator, nf = 64, is supported)");
    if (static_cast<int>(raw.at(bk).v.size()) != L.cout) return fail(c, "hdrtv_set_hg_weights: unexpected shape for " + bk);
  }
  // eval-mode BatchNorm folded into the conv (the arithmetic of Hallucination_Generator_FusedBN._fold_bn_into_conv,
  // Hallucination_arch.py:240-275, in double): scale = g * rsqrt(var + eps); w' = w * scale; b' = (b - mean) * scale + beta
  for (const char* blk : kHgBnBlocks) {
    const std::string b0 = std::string(blk) + ".0", b1 = std::string(blk) + ".1";
    if (!raw.count(b1 + ".running_var")) continue;                     // already folded (fusedbn checkpoints)
    for (const char* s : {".weight", ".bias", ".running_mean"})
      if (!raw.count(b1 + s)) return fail(c, "hdrtv_set_hg_weights: missing key " + b1 + s);
    HostTensor& w = raw.at(b0 + ".weight");
    HostTensor& b = raw.at(b0 + ".bias");
    const HostTensor &g = raw.at(b1 + ".weight"), &beta = raw.at(b1 + ".bias"), &mean = raw.at(b1 + ".running_mean"), &var = raw.at(b1 + ".running_var");
    const size_t O = static_cast<size_t>(w.shape[0]), per = w.v.size() / O;
    if (g.v.size() != O || beta.v.size() != O || mean.v.size() != O || var.v.size() != O) return fail(c, "hdrtv_set_hg_weights: BatchNorm shape for " + b1);
    for (size_t o = 0; o < O; ++o) {
      const float scale = g.v[o] * (1.0f / std::sqrt(var.v[o] + 1e-5f));
      for (size_t k = 0; k < per; ++k) w.v[o * per + k] *= scale;
      b.v[o] = (b.v[o] - mean.v[o]) * scale + beta.v[o];
    }
  }
  for (const HgLayerSpec& L : kHgLayers)
    for (const char* s : {".weight", ".bias"}) c->hg.w[std::string(L.name) + s] = std::move(raw.at(std::string(L.name) + s));
  if (c->precision == HDRTV_FP16) {
    for (const HgLayerSpec& L : kHgLayers) {
      if (L.cout == 3) continue;                                       // conv10 / conv_last live in the tail kernel
      const HgShape sh = hg_layer_shape(L);
      std::vector<__half> pk = hg_pack(c->hg.w.at(std::string(L.name) + ".weight"), c->hg.w.at(std::string(L.name) + ".bias"), sh.kind, sh.NT);
      __half* d = hg_w_upload(c, pk.data(), pk.size());
      if (!d) return fail(c, std::string("hdrtv_set_hg_weights: upload failed for ") + L.name);
      c->hg.wpk[L.name] = d;
    }
    std::unique_ptr<HgTail> tw(new HgTail());
    const HostTensor &w10 = c->hg.w.at("conv10.weight"), &b10 = c->hg.w.at("conv10.bias"), &wl = c->hg.w.at("conv_last.weight"), &bl = c->hg.w.at("conv_last.bias");
    for (int k = 0; k < 3; ++k) {
      for (int i = 0; i < 128; ++i) tw->w10[k][i] = __half2float(__float2half(w10.v[k * 128 + i]));   // the model's weights are half tensors
      for (int i = 0; i < 6; ++i) tw->wl[k][i] = __half2float(__float2half(wl.v[k * 6 + i]));
      tw->b10[k] = __half2float(__float2half(b10.v[k]));
      tw->bl[k] = __half2float(__float2half(bl.v[k]));
    }
    c->hg.d_tail = hg_w_upload(c, tw.get(), 1);
    if (!c->hg.d_tail) return fail(c, "hdrtv_set_hg_weights: upload failed for the tail");
    // conv10's weights split by producer (cat((Up_conv5 output, conv1_out), 1)): [3][64] each, for the *_DOT epilogues
    std::vector<float> dup(3 * 64), dsk(3 * 64);
    for (int k = 0; k < 3; ++k)
      for (int i = 0; i < 64; ++i) { dup[k * 64 + i] = tw->w10[k][i]; dsk[k * 64 + i] = tw->w10[k][64 + i]; }
    c->hg.d_dot_up = hg_w_upload(c, dup.data(), dup.size());
    c->hg.d_dot_skip = hg_w_upload(c, dsk.data(), dsk.size());
    if (!c->hg.d_dot_up || !c->hg.d_dot_skip) return fail(c, "hdrtv_set_hg_weights: upload failed for the conv10 split");
  } else {
    for (auto& kv : c->hg.w) {
      float* d = hg_w_upload(c, kv.second.v.data(), kv.second.v.size());
      if (!d) return fail(c, "hdrtv_set_hg_weights: upload failed for " + kv.first);
      c->hg.wd[kv.first] = d;
    }
  }
  c->hg.has = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------ FP16 plan
static int hg_add(Ctx* c, const char* layer, int epi, const P8& in0, const P8* in1, const P8& out, const P8* out_full, bool relu) {
  const HgLayerSpec* spec = nullptr;
  for (const HgLayerSpec& L : kHgLayers)
    if (std::string(L.name) == layer) spec = &L;
  if (!spec) return fail(c, std::string("hg plan: unknown layer ") + layer);
  const HgShape sh = hg_layer_shape(*spec);
  const int gc = g_group_channels(sh.kind);
  HgLaunch L;
  memset(&L.p, 0, sizeof(L.p));
  GConvParams& p = L.p;
  p.in0 = reinterpret_cast<const uint4*>(in0.base);
  p.in0_row_entries = in0.row_entries();
  p.in0_wp = static_cast<uint32_t>(in0.Wp);
  p.in0_groups = sh.kind == G_3x3_C8 ? 1 : in0.chunks * 8 / gc;
  if (in1) {
    if (in1->H != in0.H || in1->W != in0.W || in1->Wp != in0.Wp) return fail(c, std::string("hg plan: concat operands differ for ") + layer);
    p.in1 = reinterpret_cast<const uint4*>(in1->base);
    p.in1_row_entries = in1->row_entries();
    p.in1_wp = static_cast<uint32_t>(in1->Wp);
  }
  p.kgroups = p.in0_groups + (in1 ? in1->chunks * 8 / gc : 0);
  const int cin_real = (in0.chunks + (in1 ? in1->chunks : 0)) * 8;
  if (sh.kind != G_3x3_C8 && (cin_real != spec->cin || cin_real % gc != 0)) return fail(c, std::string("hg plan: channel mismatch for ") + layer);
  p.wpk = reinterpret_cast<const uint4*>(c->hg.wpk.at(layer));
  p.w_tile_bytes = (static_cast<long>(p.kgroups) * g_steps(sh.kind) + 1) * sh.NT * 32;
  p.ntiles = (spec->cout + sh.NT - 1) / sh.NT;
  p.H = in0.H;
  p.W = in0.W;
  // two-row tiles with alternating accumulators (hg.cuh, RB): measured on every 3x3 layer with <= 8 K groups per tile - only the
  // PixelShuffle layer with 4 K groups gains (Up_conv5 135 -> 125 us at 1080p, 498 -> 479 us at 4K); conv1 / conv2 / conv3_1 lose
  // 4-15 % to the extra L2 traffic of the shorter tiles
  const bool ps = epi == GE_PS || epi == GE_PS_DOT;
  const int rb = (sh.kind == G_3x3 && ps && p.kgroups <= env_int("HDRTV_HG_RB2_MAXKG", 4)) ? 2 : kGRows;
  p.strips = (p.W + kTileM - 1) / kTileM;
  p.rowblocks = (p.H + rb - 1) / rb;
  p.tiles = p.ntiles * p.strips * p.rowblocks;
  p.relu = relu ? 1 : 0;
  p.lvl = 0;
  while ((c->hg.Hp >> p.lvl) > p.H) ++p.lvl;            // resolution level of this conv's pixels
  if ((c->hg.Hp >> p.lvl) != p.H || (c->hg.Wp >> p.lvl) != p.W) return fail(c, std::string("hg plan: level mismatch for ") + layer);
  p.out = out;
  if (out_full) { p.out_full = *out_full; p.has_full = 1; }
  p.err = c->d_err;
  const bool pool = epi == GE_POOL || epi == GE_POOL_DOT;
  const int expect_c = epi == GE_PS ? spec->cout / 4 : spec->cout;
  const int eh = pool ? p.H / 2 : (epi == GE_PS ? 2 * p.H : p.H), ew = pool ? p.W / 2 : (epi == GE_PS ? 2 * p.W : p.W);
  if (epi != GE_PS_DOT && (out.chunks * 8 != expect_c || out.H != eh || out.W != ew)) return fail(c, std::string("hg plan: output tensor mismatch for ") + layer);
  if (pool && ((p.H | p.W) & 1)) return fail(c, std::string("hg plan: pooled layer needs even dimensions: ") + layer);
  if (epi == GE_POOL_DOT || epi == GE_PS_DOT) {
    // conv10 folded into this producer: partial sums into slices [first, first + 2 * ntiles) of the context's buffer
    const bool skip = epi == GE_POOL_DOT;
    const int Hd = skip ? p.H : 2 * p.H, Wd2 = skip ? p.W : 2 * p.W;
    if (!c->hg.d_part || (skip ? spec->cout : spec->cout / 4) != 64) return fail(c, std::string("hg plan: conv10 fold needs a 64-channel producer: ") + layer);
    p.dot_w = skip ? c->hg.d_dot_skip : c->hg.d_dot_up;
    p.dot_out = c->hg.d_part + static_cast<long>(skip ? 0 : 2) * 3 * Hd * Wd2;
    p.dot_H = Hd;
    p.dot_W = Wd2;
  }
  L.kind = sh.kind;
  L.NT = sh.NT;
  L.epi = epi;
  L.rb = rb;
  L.grid = std::min(p.tiles, c->hg.sms);
  L.smem = g_smem_bytes(sh.kind, sh.NT, rb);
  L.name = layer;
  c->hg.plan.push_back(L);
  return 0;
}

template <int KIND, int NT, int EPI, int RB = kGRows>
static cudaError_t hg_launch_t(const HgLaunch& L, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gconv_kernel<KIND, NT, EPI, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(g_smem_bytes(KIND, NT, RB)));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  gconv_kernel<KIND, NT, EPI, RB><<<L.grid, kGThreads, L.smem, s>>>(L.p);
  return cudaGetLastError();
}
static cudaError_t hg_launch(const HgLaunch& L, cudaStream_t s) {
  if (L.rb == 2) {
    if (L.kind == G_3x3 && L.NT == 128 && L.epi == GE_PS) return hg_launch_t<G_3x3, 128, GE_PS, 2>(L, s);
    if (L.kind == G_3x3 && L.NT == 128 && L.epi == GE_PS_DOT) return hg_launch_t<G_3x3, 128, GE_PS_DOT, 2>(L, s);
    return cudaErrorInvalidValue;
  }
  if (L.kind == G_3x3_C8 && L.NT == 64 && L.epi == GE_POOL) return hg_launch_t<G_3x3_C8, 64, GE_POOL>(L, s);
  if (L.kind == G_3x3_C8 && L.NT == 64 && L.epi == GE_POOL_DOT) return hg_launch_t<G_3x3_C8, 64, GE_POOL_DOT>(L, s);
  if (L.kind == G_3x3 && L.NT == 128 && L.epi == GE_PS_DOT) return hg_launch_t<G_3x3, 128, GE_PS_DOT>(L, s);
  if (L.kind == G_3x3 && L.NT == 128 && L.epi == GE_P8) return hg_launch_t<G_3x3, 128, GE_P8>(L, s);
  if (L.kind == G_3x3 && L.NT == 128 && L.epi == GE_POOL) return hg_launch_t<G_3x3, 128, GE_POOL>(L, s);
  if (L.kind == G_3x3 && L.NT == 128 && L.epi == GE_PS) return hg_launch_t<G_3x3, 128, GE_PS>(L, s);
  if (L.kind == G_1x1 && L.NT == 128 && L.epi == GE_P8) return hg_launch_t<G_1x1, 128, GE_P8>(L, s);
  if (L.kind == G_1x1 && L.NT == 64 && L.epi == GE_P8) return hg_launch_t<G_1x1, 64, GE_P8>(L, s);
  return cudaErrorInvalidValue;
}

static int hg_prepare(Ctx* c, int H, int Wd) {
  if (!c->hg.has) return fail(c, "hdrtv_hg: HG weights not set (hdrtv_set_hg_weights)");
  if (c->hg.H == H && c->hg.W == Wd) return 0;
  if (H < 16 || Wd < 16) return fail(c, "hdrtv_hg: frame must be at least 16x16");
  cudaDeviceSynchronize();
  hg_release_ws(c);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, c->device);
  c->hg.sms = prop.multiProcessorCount;
  // HG_Composite_arch.py:91-93: five 2x poolings -> pad to the next multiple of 32
  const int Hp = rup(H, 32), Wp = rup(Wd, 32);
  // F.pad(mode="reflect") needs pad < size (torch raises otherwise): a 16-row frame cannot be reflected up to 32 rows
  if (Hp - H >= H || Wp - Wd >= Wd)
    return fail(c, "hdrtv_hg: reflect padding to a multiple of 32 needs more rows / columns than the padding (" + std::to_string(H) + "x" +
                       std::to_string(Wd) + "), as torch.nn.functional.pad(mode='reflect') does");
  c->hg.Hp = Hp;
  c->hg.Wp = Wp;
  int h[6], w[6];
  for (int l = 0; l < 6; ++l) { h[l] = Hp >> l; w[l] = Wp >> l; }
  if (c->precision == HDRTV_FP16) {
    auto& T = c->hg.t;
    auto mk = [&](const char* n, int C, int l) { T[n] = hg_make_p8(c, C, h[l], w[l]); return T[n].base != nullptr; };
    // conv10 (1x1 on cat(Up_conv5 output, conv1_out)) folded into its two producers' epilogues: neither 64-channel
    // full-resolution tensor reaches HBM (2 x 128 B/px written + read), only 6 slices of 3 partial sums (72 B/px)
    const bool fuse = env_int("HDRTV_HG_FUSE_CONV10", 1) != 0;
    c->hg.fuse_conv10 = fuse;
    c->hg.d_part = fuse ? hg_ws_alloc<float>(c, static_cast<size_t>(6) * 3 * Hp * Wp, true) : nullptr;
    c->hg.gate_cw = (Wp + kHgCell - 1) / kHgCell;
    c->hg.gate_ch = (Hp + kHgCell - 1) / kHgCell;
    // gate: one flag word (+ 3 spare), then the `seen` and `active` cell maps (one byte per cell each)
    c->hg.d_gate = reinterpret_cast<int*>(hg_ws_alloc<uint8_t>(c, 16 + 2 * static_cast<size_t>(c->hg.gate_cw) * c->hg.gate_ch, true));
    if ((fuse && !c->hg.d_part) || !c->hg.d_gate) return fail(c, "hdrtv_hg: workspace allocation failed");
    bool ok = mk("img", 8, 0) && (fuse || mk("c1", 64, 0)) && mk("p1", 64, 1) && mk("c2", 128, 1) && mk("p31", 256, 2) && mk("c3", 256, 2) &&
              mk("p41", 512, 3) && mk("c4", 512, 3) && mk("p51", 512, 4) && mk("c5", 512, 4) && mk("pc1", 512, 5) && mk("code", 512, 5) &&
              mk("u1", 512, 4) && mk("c6", 512, 4) && mk("u2", 512, 3) && mk("c7", 256, 3) && mk("u3", 256, 2) && mk("c8", 128, 2) &&
              mk("u4", 128, 1) && mk("c9", 64, 1) && (fuse || mk("u5", 64, 0));
    if (!ok) return fail(c, "hdrtv_hg: workspace allocation failed");
    int r = 0;
    if (fuse) r |= hg_add(c, "conv1.0", GE_POOL_DOT, T["img"], nullptr, T["p1"], nullptr, true);
    else r |= hg_add(c, "conv1.0", GE_POOL, T["img"], nullptr, T["p1"], &T["c1"], true);
    r |= hg_add(c, "conv2.0", GE_P8, T["p1"], nullptr, T["c2"], nullptr, true);
    r |= hg_add(c, "conv3_1.0", GE_POOL, T["c2"], nullptr, T["p31"], nullptr, true);
    r |= hg_add(c, "conv3_2.0", GE_P8, T["p31"], nullptr, T["c3"], nullptr, true);
    r |= hg_add(c, "conv4_1.0", GE_POOL, T["c3"], nullptr, T["p41"], nullptr, true);
    r |= hg_add(c, "conv4_2.0", GE_P8, T["p41"], nullptr, T["c4"], nullptr, true);
    r |= hg_add(c, "conv5_1.0", GE_POOL, T["c4"], nullptr, T["p51"], nullptr, true);
    r |= hg_add(c, "conv5_2.0", GE_P8, T["p51"], nullptr, T["c5"], nullptr, true);
    r |= hg_add(c, "conv_code1.0", GE_POOL, T["c5"], nullptr, T["pc1"], nullptr, true);
    r |= hg_add(c, "conv_code2.0", GE_P8, T["pc1"], nullptr, T["code"], nullptr, true);
    r |= hg_add(c, "Up_conv1.0", GE_PS, T["code"], nullptr, T["u1"], nullptr, true);
    r |= hg_add(c, "conv6", GE_P8, T["u1"], &T["c5"], T["c6"], nullptr, false);
    r |= hg_add(c, "Up_conv2.0", GE_PS, T["c6"], nullptr, T["u2"], nullptr, true);
    r |= hg_add(c, "conv7", GE_P8, T["u2"], &T["c4"], T["c7"], nullptr, false);
    r |= hg_add(c, "Up_conv3.0", GE_PS, T["c7"], nullptr, T["u3"], nullptr, true);
    r |= hg_add(c, "conv8", GE_P8, T["u3"], &T["c3"], T["c8"], nullptr, false);
    r |= hg_add(c, "Up_conv4.0", GE_PS, T["c8"], nullptr, T["u4"], nullptr, true);
    r |= hg_add(c, "conv9", GE_P8, T["u4"], &T["c2"], T["c9"], nullptr, false);
    if (fuse) r |= hg_add(c, "Up_conv5.0", GE_PS_DOT, T["c9"], nullptr, T["c9"], nullptr, true);
    else r |= hg_add(c, "Up_conv5.0", GE_PS, T["c9"], nullptr, T["u5"], nullptr, true);
    if (r) return -1;
  } else {
    auto& B = c->hg.f32;
    auto A = [&](const char* n, int C, int l) { B[n] = hg_ws_alloc<float>(c, static_cast<size_t>(C) * h[l] * w[l], false); return B[n] != nullptr; };
    bool ok = A("img", 3, 0) && A("c1", 64, 0) && A("p1", 64, 1) && A("c2", 128, 1) && A("t1", 256, 1) && A("p31", 256, 2) && A("c3", 256, 2) &&
              A("t2", 512, 2) && A("p41", 512, 3) && A("c4", 512, 3) && A("t3", 512, 3) && A("p51", 512, 4) && A("c5", 512, 4) &&
              A("t4", 512, 4) && A("pc1", 512, 5) && A("code", 512, 5) && A("u1", 512, 4) && A("c6", 512, 4) && A("u2", 512, 3) &&
              A("c7", 256, 3) && A("u3", 256, 2) && A("c8", 128, 2) && A("u4", 128, 1) && A("c9", 64, 1) && A("u5", 64, 0) &&
              A("c10", 3, 0) && A("o", 3, 0);
    if (!ok) return fail(c, "hdrtv_hg: fp32 workspace allocation failed");
  }
  CK(c, cudaDeviceSynchronize());
  c->hg.H = H;
  c->hg.W = Wd;
  return 0;
}

static int hg_conv32(Ctx* c, cudaStream_t s, const char* layer, const float* in0, int C0, const float* in1, int C1, float* out, int H, int Wd,
                     bool relu, bool ps) {
  const HostTensor& w = c->hg.w.at(std::string(layer) + ".weight");
  HgConvF32 p;
  p.in0 = in0; p.C0 = C0; p.in1 = in1; p.C1 = C1;
  p.w = c->hg.wd.at(std::string(layer) + ".weight");
  p.b = c->hg.wd.at(std::string(layer) + ".bias");
  p.out = out;
  p.Cout = static_cast<int>(w.shape[0]);
  p.H = H; p.W = Wd;
  p.ks = static_cast<int>(w.shape[2]);
  p.relu = relu ? 1 : 0;
  p.ps = ps ? 1 : 0;
  if (static_cast<int>(w.shape[1]) != C0 + C1) return fail(c, std::string("hg fp32: channel mismatch for ") + layer);
  constexpr int PXT = 4;
  if (p.Cout >= 16) {
    dim3 grid((Wd + 64 * PXT - 1) / (64 * PXT), H, (p.Cout + 15) / 16);
    hg_conv_f32_kernel<16, PXT><<<grid, 64, sizeof(float) * kHgF32Chunk * p.ks * p.ks * 16, s>>>(p);
  } else {
    dim3 grid((Wd + 64 * PXT - 1) / (64 * PXT), H, (p.Cout + 7) / 8);
    hg_conv_f32_kernel<8, PXT><<<grid, 64, sizeof(float) * kHgF32Chunk * p.ks * p.ks * 8, s>>>(p);
  }
  CK(c, cudaGetLastError());
  ++c->launches;
  return 0;
}

// base_out: planar (3, H, W) in the context's precision; out: planar fp32 (3, H, W) (the reference's HG output is a
// float tensor in both precisions: mask.float() * out + img promotes, HG_Composite_arch.py:83, Hallucination_arch.py:136)
static int hg_run(Ctx* c, const void* base_out, int H, int Wd, float* out, cudaStream_t s) {
  if (hg_prepare(c, H, Wd)) return -1;
  const int Hp = c->hg.Hp, Wp = c->hg.Wp;
  if (c->precision == HDRTV_FP16) {
    auto& T = c->hg.t;
    // Highlight gate.  The stage's output differs from its input only inside the mask (max_c(base) > 0.775: speculars,
    // lamps, white areas), and a masked output depends on nothing further than 186 pixels away.  The stage-in pass marks
    // the 64 x 64 cells that hold a masked pixel, a one-block kernel grows that map by three cells, and the 19 U-Net
    // launches compute only the tiles that touch an active cell (none at all: they return at once).  No host
    // synchronisation, the launch sequence is the same for every frame.  Bit-identical to the dense evaluation; HDRTV_HG_EARLY_OUT=0 always runs
    // the whole U-Net (benchmarks quote that figure).  Needs the conv10 fold's tail kernel.
    const char* eo_env = getenv("HDRTV_HG_EARLY_OUT");
    const bool early = c->hg.fuse_conv10 && c->hg.d_gate && !(eo_env && eo_env[0] == '0');
    int* gate = early ? c->hg.d_gate : nullptr;
    const int cells = c->hg.gate_cw * c->hg.gate_ch;
    uint8_t* seen = gate ? reinterpret_cast<uint8_t*>(gate) + 16 : nullptr;
    uint8_t* active = gate ? seen + cells : nullptr;
    if (gate) {
      CK(c, cudaMemsetAsync(gate, 0x7f, 16, s));                           // kHgNoMask
      CK(c, cudaMemsetAsync(seen, 0, cells, s));
    }
    hg_stage_in_kernel<__half><<<dim3((Wp + 127) / 128, Hp), 128, 0, s>>>(static_cast<const __half*>(base_out), T.at("img"), H, Wd, Hp, Wp, gate,
                                                                          seen, c->hg.gate_cw);
    CK(c, cudaGetLastError());
    ++c->launches;
    if (gate) {
      hg_gate_dilate_kernel<<<1, 256, 0, s>>>(seen, active, c->hg.gate_cw, c->hg.gate_ch);
      CK(c, cudaGetLastError());
      ++c->launches;
    }
    for (HgLaunch& L : c->hg.plan) {
      L.p.gate = gate;
      L.p.gate_active = active;
      L.p.gate_cw = c->hg.gate_cw;
      L.p.gate_ch = c->hg.gate_ch;
      CK(c, hg_launch(L, s));
      ++c->launches;
    }
    if (c->hg.fuse_conv10)
      hg_tail_dot_kernel<<<dim3((Wd + 127) / 128, H), 128, 0, s>>>(c->hg.d_part, 6, Hp, Wp, T.at("img"), c->hg.d_tail, out, H, Wd, gate);
    else
      hg_tail_kernel<<<dim3((Wd + 127) / 128, H), 128, 0, s>>>(T.at("u5"), T.at("c1"), T.at("img"), c->hg.d_tail, out, H, Wd);
    CK(c, cudaGetLastError());
    ++c->launches;
    return 0;
  }
  auto B = [&](const char* n) { return c->hg.f32.at(n); };
  int h[6], w[6];
  for (int l = 0; l < 6; ++l) { h[l] = Hp >> l; w[l] = Wp >> l; }
  hg_reflect_f32_kernel<<<dim3((Wp + 127) / 128, Hp, 3), 128, 0, s>>>(static_cast<const float*>(base_out), B("img"), H, Wd, Hp, Wp);
  CK(c, cudaGetLastError());
  ++c->launches;
  auto pool = [&](const float* in, float* o, int C, int l) {
    const long n = static_cast<long>(C) * h[l + 1] * w[l + 1];
    hg_maxpool_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(in, o, C, h[l], w[l]);
    ++c->launches;
  };
  int r = 0;
  r |= hg_conv32(c, s, "conv1.0", B("img"), 3, nullptr, 0, B("c1"), h[0], w[0], true, false);
  pool(B("c1"), B("p1"), 64, 0);
  r |= hg_conv32(c, s, "conv2.0", B("p1"), 64, nullptr, 0, B("c2"), h[1], w[1], true, false);
  r |= hg_conv32(c, s, "conv3_1.0", B("c2"), 128, nullptr, 0, B("t1"), h[1], w[1], true, false);
  pool(B("t1"), B("p31"), 256, 1);
  r |= hg_conv32(c, s, "conv3_2.0", B("p31"), 256, nullptr, 0, B("c3"), h[2], w[2], true, false);
  r |= hg_conv32(c, s, "conv4_1.0", B("c3"), 256, nullptr, 0, B("t2"), h[2], w[2], true, false);
  pool(B("t2"), B("p41"), 512, 2);
  r |= hg_conv32(c, s, "conv4_2.0", B("p41"), 512, nullptr, 0, B("c4"), h[3], w[3], true, false);
  r |= hg_conv32(c, s, "conv5_1.0", B("c4"), 512, nullptr, 0, B("t3"), h[3], w[3], true, false);
  pool(B("t3"), B("p51"), 512, 3);
  r |= hg_conv32(c, s, "conv5_2.0", B("p51"), 512, nullptr, 0, B("c5"), h[4], w[4], true, false);
  r |= hg_conv32(c, s, "conv_code1.0", B("c5"), 512, nullptr, 0, B("t4"), h[4], w[4], true, false);
  pool(B("t4"), B("pc1"), 512, 4);
  r |= hg_conv32(c, s, "conv_code2.0", B("pc1"), 512, nullptr, 0, B("code"), h[5], w[5], true, false);
  r |= hg_conv32(c, s, "Up_conv1.0", B("code"), 512, nullptr, 0, B("u1"), h[5], w[5], true, true);
  r |= hg_conv32(c, s, "conv6", B("u1"), 512, B("c5"), 512, B("c6"), h[4], w[4], false, false);
  r |= hg_conv32(c, s, "Up_conv2.0", B("c6"), 512, nullptr, 0, B("u2"), h[4], w[4], true, true);
  r |= hg_conv32(c, s, "conv7", B("u2"), 512, B("c4"), 512, B("c7"), h[3], w[3], false, false);
  r |= hg_conv32(c, s, "Up_conv3.0", B("c7"), 256, nullptr, 0, B("u3"), h[3], w[3], true, true);
  r |= hg_conv32(c, s, "conv8", B("u3"), 256, B("c3"), 256, B("c8"), h[2], w[2], false, false);
  r |= hg_conv32(c, s, "Up_conv4.0", B("c8"), 128, nullptr, 0, B("u4"), h[2], w[2], true, true);
  r |= hg_conv32(c, s, "conv9", B("u4"), 128, B("c2"), 128, B("c9"), h[1], w[1], false, false);
  r |= hg_conv32(c, s, "Up_conv5.0", B("c9"), 64, nullptr, 0, B("u5"), h[1], w[1], true, true);
  r |= hg_conv32(c, s, "conv10", B("u5"), 64, B("c1"), 64, B("c10"), h[0], w[0], false, false);
  r |= hg_conv32(c, s, "conv_last", B("c10"), 3, B("img"), 3, B("o"), h[0], w[0], false, false);
  if (r) return -1;
  hg_blend_f32_kernel<<<dim3((Wd + 127) / 128, H), 128, 0, s>>>(B("o"), B("img"), out, H, Wd, Hp, Wp);
  CK(c, cudaGetLastError());
  ++c->launches;
  return 0;
}

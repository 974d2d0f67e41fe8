// Host side of the GPU letterbox (gui_scaling.py:228-244): geometry, OpenCV's resize tables (resize.cpp:
// computeResizeAreaTab, interpolateCubic + fixed-point weights), launch.  Included by engine.cu inside namespace hdrtv.
#pragma once

struct LbAxisArea { std::vector<int> ofs, cnt; std::vector<float> w; int K = 0; };
static LbAxisArea lb_area_axis(int ssize, int dsize) {
  const double scale = 1.0 / (static_cast<double>(dsize) / static_cast<double>(ssize));
  struct E { int d, s; float a; };
  std::vector<E> tab;
  for (int dx = 0; dx < dsize; ++dx) {
    const double fsx1 = dx * scale, fsx2 = fsx1 + scale, cell = std::min(scale, ssize - fsx1);
    int sx1 = static_cast<int>(std::ceil(fsx1)), sx2 = static_cast<int>(std::floor(fsx2));
    sx2 = std::min(sx2, ssize - 1);
    sx1 = std::min(sx1, sx2);
    if (sx1 - fsx1 > 1e-3) tab.push_back({dx, sx1 - 1, static_cast<float>((sx1 - fsx1) / cell)});
    for (int sx = sx1; sx < sx2; ++sx) tab.push_back({dx, sx, static_cast<float>(1.0 / cell)});
    if (fsx2 - sx2 > 1e-3) tab.push_back({dx, sx2, static_cast<float>(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell)});
  }
  LbAxisArea ax;
  ax.ofs.assign(dsize, 0);
  ax.cnt.assign(dsize, 0);
  for (const E& e : tab) {
    if (ax.cnt[e.d] == 0) ax.ofs[e.d] = e.s;
    ++ax.cnt[e.d];
  }
  for (int c : ax.cnt) ax.K = std::max(ax.K, c);
  ax.w.assign(static_cast<size_t>(dsize) * ax.K, 0.f);
  std::vector<int> fill(dsize, 0);
  for (const E& e : tab) ax.w[static_cast<size_t>(e.d) * ax.K + fill[e.d]++] = e.a;
  return ax;
}
struct LbAxisCubic { std::vector<int> idx, ia; std::vector<float> fb; };
static LbAxisCubic lb_cubic_axis(int ssize, int dsize) {
  const double scale = 1.0 / (static_cast<double>(dsize) / static_cast<double>(ssize));
  LbAxisCubic ax;
  ax.idx.resize(static_cast<size_t>(dsize) * 4);
  ax.ia.resize(static_cast<size_t>(dsize) * 4);
  ax.fb.resize(static_cast<size_t>(dsize) * 4);
  for (int d = 0; d < dsize; ++d) {
    volatile float f = static_cast<float>((d + 0.5) * scale - 0.5);
    const int s = static_cast<int>(std::floor(f));
    volatile float x = f - static_cast<float>(s);
    // interpolateCubic, every operation rounded to fp32 (volatile temporaries: no contraction, no excess precision)
    const float A = -0.75f;
    volatile float x1 = x + 1.f;
    volatile float t = A * x1; t = t - 5.f * A; t = t * x1; t = t + 8.f * A; t = t * x1; t = t - 4.f * A;
    const float c0 = t;
    volatile float u = (A + 2.f) * x; u = u - (A + 3.f); u = u * x; u = u * x; u = u + 1.f;
    const float c1 = u;
    volatile float y = 1.f - x;
    volatile float v = (A + 2.f) * y; v = v - (A + 3.f); v = v * y; v = v * y; v = v + 1.f;
    const float c2 = v;
    volatile float w3 = 1.f - c0; w3 = w3 - c1; w3 = w3 - c2;
    const float c[4] = {c0, c1, c2, w3};
    for (int k = 0; k < 4; ++k) {
      volatile float sc = c[k] * 2048.f;
      const int ia = static_cast<int>(std::nearbyint(static_cast<float>(sc)));       // saturate_cast<short>: round-half-even
      ax.idx[d * 4 + k] = std::min(std::max(s - 1 + k, 0), ssize - 1);
      ax.ia[d * 4 + k] = ia;
      ax.fb[d * 4 + k] = static_cast<float>(ia) * (1.f / (2048.f * 2048.f));
    }
  }
  return ax;
}

template <typename T>
static const T* lb_upload(Ctx* c, const std::vector<T>& v) {
  void* p = nullptr;
  if (v.empty() || cudaMalloc(&p, v.size() * sizeof(T)) != cudaSuccess) return nullptr;
  cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  c->lb.allocs.push_back(p);
  return static_cast<const T*>(p);
}
static void lb_release(Ctx* c) {
  for (void* p : c->lb.allocs) cudaFree(p);
  c->lb.allocs.clear();
  c->lb.key[0] = c->lb.key[1] = c->lb.key[2] = c->lb.key[3] = 0;
}

static int lb_prepare(Ctx* c, int H, int Wd, int out_H, int out_W) {
  LetterboxState& L = c->lb;
  if (L.key[0] == H && L.key[1] == Wd && L.key[2] == out_H && L.key[3] == out_W) return 0;
  cudaDeviceSynchronize();
  lb_release(c);
  Letterbox& p = L.p;
  memset(&p, 0, sizeof(p));
  // gui_scaling.py:234-237 (python floats are doubles; round() is round-half-even = nearbyint in the default mode)
  const double scale = std::min(static_cast<double>(out_W) / std::max(Wd, 1), static_cast<double>(out_H) / std::max(H, 1));
  p.new_W = std::max(1, static_cast<int>(std::nearbyint(Wd * scale)));
  p.new_H = std::max(1, static_cast<int>(std::nearbyint(H * scale)));
  if (p.new_W > out_W || p.new_H > out_H) return fail(c, "hdrtv_letterbox_bgr: resized frame exceeds the canvas");
  p.H = H; p.W = Wd; p.out_H = out_H; p.out_W = out_W;
  p.x0 = (out_W - p.new_W) / 2;
  p.y0 = (out_H - p.new_H) / 2;
  if (p.new_W == Wd && p.new_H == H) p.mode = LB_COPY;
  else if (scale < 1.0 && H % p.new_H == 0 && Wd % p.new_W == 0) {
    p.mode = LB_AREA_INT;
    p.fy = H / p.new_H;
    p.fx = Wd / p.new_W;
  } else if (scale < 1.0) {
    p.mode = LB_AREA;
    const LbAxisArea ax = lb_area_axis(Wd, p.new_W), ay = lb_area_axis(H, p.new_H);
    p.xofs = lb_upload(c, ax.ofs); p.xcnt = lb_upload(c, ax.cnt); p.xw = lb_upload(c, ax.w); p.KX = ax.K;
    p.yofs = lb_upload(c, ay.ofs); p.ycnt = lb_upload(c, ay.cnt); p.yw = lb_upload(c, ay.w); p.KY = ay.K;
    if (!p.xofs || !p.xcnt || !p.xw || !p.yofs || !p.ycnt || !p.yw) return fail(c, "hdrtv_letterbox_bgr: table upload failed");
  } else {
    p.mode = LB_CUBIC;
    const LbAxisCubic ax = lb_cubic_axis(Wd, p.new_W), ay = lb_cubic_axis(H, p.new_H);
    p.xofs = lb_upload(c, ax.idx); p.xi = lb_upload(c, ax.ia);
    p.yofs = lb_upload(c, ay.idx); p.yi = lb_upload(c, ay.ia); p.yw = lb_upload(c, ay.fb);
    if (!p.xofs || !p.xi || !p.yofs || !p.yi || !p.yw) return fail(c, "hdrtv_letterbox_bgr: table upload failed");
  }
  L.key[0] = H; L.key[1] = Wd; L.key[2] = out_H; L.key[3] = out_W;
  return 0;
}

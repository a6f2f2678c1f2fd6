/* ORACLE — TEST INFRASTRUCTURE ONLY (not part of the product path).
 *
 * Plain-C, direct-loop, double-accumulating restatement of the primitive operators the reference's
 * hot path issues through torch.nn (model.py / losses.py of /root/reference). It exists so that the
 * torch-based oracle (oracle/oracle.py) is itself pinned by an implementation that shares no code
 * with PyTorch: tests/test_oracle_c.py checks every function here against torch on small shapes.
 * Tensors are float32, NCHW, contiguous.
 *
 * gcc -O2 -shared -fPIC -o _build/liboracle.so msig_oracle.c -lm
 */
#include <math.h>
#include <stddef.h>

static int reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

/* nn.Conv2d(cin, cout, k, stride, pad, padding_mode = zeros | reflect): model.py:45,48,72-75,131-133,141,165,183 */
void orc_conv2d(const float* x, const float* w, const float* b, float* y, int n, int cin, int h, int wd, int cout,
                int k, int stride, int pad_t, int pad_l, int oh, int ow, int reflect_pad) {
  for (int in_ = 0; in_ < n; ++in_)
    for (int co = 0; co < cout; ++co)
      for (int yo = 0; yo < oh; ++yo)
        for (int xo = 0; xo < ow; ++xo) {
          double acc = b ? b[co] : 0.0;
          for (int ci = 0; ci < cin; ++ci)
            for (int r = 0; r < k; ++r)
              for (int s = 0; s < k; ++s) {
                int iy = yo * stride + r - pad_t, ix = xo * stride + s - pad_l;
                if (reflect_pad) {
                  iy = reflect(iy, h);
                  ix = reflect(ix, wd);
                } else if (iy < 0 || iy >= h || ix < 0 || ix >= wd) {
                  continue;
                }
                acc += (double)x[((size_t)(in_ * cin + ci) * h + iy) * wd + ix] *
                       (double)w[((size_t)(co * cin + ci) * k + r) * k + s];
              }
          y[((size_t)(in_ * cout + co) * oh + yo) * ow + xo] = (float)acc;
        }
}

/* gradients of orc_conv2d w.r.t. x, w, b given dy (any of dx/dw/db may be NULL) */
void orc_conv2d_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int n, int cin,
                    int h, int wd, int cout, int k, int stride, int pad_t, int pad_l, int oh, int ow, int reflect_pad) {
  if (dx) for (size_t i = 0; i < (size_t)n * cin * h * wd; ++i) dx[i] = 0.f;
  if (dw) for (size_t i = 0; i < (size_t)cout * cin * k * k; ++i) dw[i] = 0.f;
  if (db) for (int i = 0; i < cout; ++i) db[i] = 0.f;
  for (int in_ = 0; in_ < n; ++in_)
    for (int co = 0; co < cout; ++co)
      for (int yo = 0; yo < oh; ++yo)
        for (int xo = 0; xo < ow; ++xo) {
          const float g = dy[((size_t)(in_ * cout + co) * oh + yo) * ow + xo];
          if (db) db[co] += g;
          for (int ci = 0; ci < cin; ++ci)
            for (int r = 0; r < k; ++r)
              for (int s = 0; s < k; ++s) {
                int iy = yo * stride + r - pad_t, ix = xo * stride + s - pad_l;
                if (reflect_pad) {
                  iy = reflect(iy, h);
                  ix = reflect(ix, wd);
                } else if (iy < 0 || iy >= h || ix < 0 || ix >= wd) {
                  continue;
                }
                const size_t xi = ((size_t)(in_ * cin + ci) * h + iy) * wd + ix;
                const size_t wi = ((size_t)(co * cin + ci) * k + r) * k + s;
                if (dx) dx[xi] += g * w[wi];
                if (dw) dw[wi] += g * x[xi];
              }
        }
}

/* nn.ConvTranspose2d(cin, cout, 4, 2, 1), weight [cin][cout][4][4]: model.py:139-140 */
void orc_conv_transpose2d(const float* x, const float* w, const float* b, float* y, int n, int cin, int h, int wd,
                          int cout) {
  const int oh = 2 * h, ow = 2 * wd;
  for (size_t i = 0; i < (size_t)n * cout * oh * ow; ++i) y[i] = 0.f;
  for (int in_ = 0; in_ < n; ++in_)
    for (int ci = 0; ci < cin; ++ci)
      for (int iy = 0; iy < h; ++iy)
        for (int ix = 0; ix < wd; ++ix) {
          const float v = x[((size_t)(in_ * cin + ci) * h + iy) * wd + ix];
          for (int co = 0; co < cout; ++co)
            for (int r = 0; r < 4; ++r)
              for (int s = 0; s < 4; ++s) {
                const int yo = 2 * iy - 1 + r, xo = 2 * ix - 1 + s;
                if (yo < 0 || yo >= oh || xo < 0 || xo >= ow) continue;
                y[((size_t)(in_ * cout + co) * oh + yo) * ow + xo] += v * w[((size_t)(ci * cout + co) * 4 + r) * 4 + s];
              }
        }
  if (b)
    for (int in_ = 0; in_ < n; ++in_)
      for (int co = 0; co < cout; ++co)
        for (int p = 0; p < oh * ow; ++p) y[(size_t)(in_ * cout + co) * oh * ow + p] += b[co];
}

/* nn.InstanceNorm2d(affine=False, eps) followed by y = gamma*xhat + beta (AdaIN, model.py:20-36);
 * gamma/beta [n][c] or NULL (plain InstanceNorm). Biased variance, eps inside the sqrt. */
void orc_adain(const float* x, const float* gamma, const float* beta, float* y, int n, int c, int hw, float eps) {
  for (int i = 0; i < n * c; ++i) {
    const float* xp = x + (size_t)i * hw;
    double m = 0.0, v = 0.0;
    for (int p = 0; p < hw; ++p) m += xp[p];
    m /= hw;
    for (int p = 0; p < hw; ++p) v += (xp[p] - m) * (xp[p] - m);
    v /= hw;
    const double r = 1.0 / sqrt(v + eps);
    const double g = gamma ? gamma[i] : 1.0, b = beta ? beta[i] : 0.0;
    for (int p = 0; p < hw; ++p) y[(size_t)i * hw + p] = (float)(g * (xp[p] - m) * r + b);
  }
}

/* backward of orc_adain: dx, dgamma [n][c], dbeta [n][c] */
void orc_adain_bwd(const float* x, const float* gamma, const float* dy, float* dx, float* dgamma, float* dbeta, int n,
                   int c, int hw, float eps) {
  for (int i = 0; i < n * c; ++i) {
    const float* xp = x + (size_t)i * hw;
    const float* gp = dy + (size_t)i * hw;
    double m = 0.0, v = 0.0;
    for (int p = 0; p < hw; ++p) m += xp[p];
    m /= hw;
    for (int p = 0; p < hw; ++p) v += (xp[p] - m) * (xp[p] - m);
    v /= hw;
    const double r = 1.0 / sqrt(v + eps);
    const double g = gamma ? gamma[i] : 1.0;
    double s1 = 0.0, s2 = 0.0;
    for (int p = 0; p < hw; ++p) {
      s1 += gp[p];
      s2 += gp[p] * (xp[p] - m) * r;
    }
    if (dgamma) dgamma[i] = (float)s2;
    if (dbeta) dbeta[i] = (float)s1;
    for (int p = 0; p < hw; ++p) {
      const double xh = (xp[p] - m) * r;
      dx[(size_t)i * hw + p] = (float)(g * r * (gp[p] - s1 / hw - xh * s2 / hw));
    }
  }
}

/* compute_gram_matrix (losses.py:70-78): x [a][b][c*d] -> G [a*b][a*b] = F F^T / (a*b*c*d) */
void orc_gram(const float* x, float* g, int ab, int cd) {
  for (int i = 0; i < ab; ++i)
    for (int j = 0; j < ab; ++j) {
      double acc = 0.0;
      for (int p = 0; p < cd; ++p) acc += (double)x[(size_t)i * cd + p] * (double)x[(size_t)j * cd + p];
      g[(size_t)i * ab + j] = (float)(acc / ((double)ab * cd));
    }
}

/* nn.L1Loss / F.l1_loss (mean) and nn.MSELoss (mean): trainer.py:50-52, losses.py:88,97 */
float orc_l1(const float* a, const float* b, size_t n) {
  double s = 0.0;
  for (size_t i = 0; i < n; ++i) s += fabs((double)a[i] - (double)b[i]);
  return (float)(s / (double)n);
}
float orc_mse(const float* a, const float* b, size_t n) {
  double s = 0.0;
  for (size_t i = 0; i < n; ++i) s += ((double)a[i] - (double)b[i]) * ((double)a[i] - (double)b[i]);
  return (float)(s / (double)n);
}

/* nn.MaxPool2d(2) (VGG pool_2 / pool_4) and AdaptiveAvgPool2d(1) (model.py:76) */
void orc_maxpool2(const float* x, float* y, int nc, int h, int w) {
  for (int i = 0; i < nc; ++i)
    for (int yo = 0; yo < h / 2; ++yo)
      for (int xo = 0; xo < w / 2; ++xo) {
        const float* p = x + ((size_t)i * h + 2 * yo) * w + 2 * xo;
        float m = p[0];
        if (p[1] > m) m = p[1];
        if (p[w] > m) m = p[w];
        if (p[w + 1] > m) m = p[w + 1];
        y[((size_t)i * (h / 2) + yo) * (w / 2) + xo] = m;
      }
}
void orc_avgpool(const float* x, float* y, int nc, int hw) {
  for (int i = 0; i < nc; ++i) {
    double s = 0.0;
    for (int p = 0; p < hw; ++p) s += x[(size_t)i * hw + p];
    y[i] = (float)(s / hw);
  }
}

// Implicit reparameterisation gradients of Gamma and Beta draws,  dz/dalpha = -(dCDF/dalpha)/pdf.
//
// torch.distributions' Gamma.rsample / Beta.rsample (TORCH distributions/gamma.py:79-87,
// dirichlet.py:16-35) back-propagate through `_standard_gamma_grad` / `_dirichlet_grad`, whose
// scalar algorithms live in TORCH include/ATen/native/Distributions.h:303-369 (gamma) and :374-511
// (beta / dirichlet). They are piecewise approximations (Taylor series near 0, Rice saddle-point
// expansion for large shape, fitted rational function in between) from Jankowiak & Obermeyer,
// "Pathwise Derivatives Beyond the Reparameterization Trick" (ICML 2018). Matching the reference
// to fp32 tolerance means using the same regimes and the same fitted coefficients; the code below
// restates them in fp64 (ATen's CPU path uses accscalar_t = double for float inputs).
#pragma once

#include "common.cuh"

namespace mnf {

// d/dalpha of a standard Gamma(alpha, 1) draw x.   ATen Distributions.h:303-369
static __device__ __noinline__ double standard_gamma_grad(double alpha, double x) {
  if (x < 0.8) {
    // Taylor series of the lower incomplete gamma function around x = 0.
    double term = 1.0, a = alpha;
    double s1 = 1.0 / a, s2 = 1.0 / (a * a);
    for (int i = 1; i <= 5; ++i) {
      term *= -x / (double)i;
      a += 1.0;
      s1 += term / a;
      s2 += term / (a * a);
    }
    const double xa = pow(x, alpha);
    const double pdf = pow(x, alpha - 1.0) * exp(-x);
    const double cdf = xa * s1;
    const double dcdf = (log(x) - digamma_d(alpha)) * cdf - xa * s2;
    const double g = -dcdf / pdf;
    return g != g ? 0.0 : g;
  }
  if (alpha > 8.0) {
    // Rice saddle-point expansion; a polynomial patch removes the singularity at x == alpha.
    if (0.9 * alpha <= x && x <= 1.1 * alpha) {
      const double n1 = 1.0 + 24.0 * alpha * (1.0 + 12.0 * alpha);
      const double n2 = 1440.0 * (alpha * alpha) + 6.0 * x * (53.0 - 120.0 * x) -
                        65.0 * x * x / alpha + alpha * (107.0 + 3600.0 * x);
      const double den = 1244160.0 * (alpha * alpha) * (alpha * alpha);
      return n1 * n2 / den;
    }
    const double den = sqrt(8.0 * alpha);
    const double t2 = den / (alpha - x);
    const double t3 = pow(x - alpha - alpha * log(x / alpha), -1.5);
    const double t23 = (x < alpha) ? t2 - t3 : t2 + t3;
    const double t1 = log(x / alpha) * t23 -
                      sqrt(2.0 / alpha) * (alpha + x) / ((alpha - x) * (alpha - x));
    const double stirling = 1.0 + 1.0 / (12.0 * alpha) * (1.0 + 1.0 / (24.0 * alpha));
    return -stirling * (x * t1) / den;
  }
  // Fitted bivariate rational function in u = log(x/alpha), v = log(alpha).
  const double u = log(x / alpha);
  const double v = log(alpha);
  const double K[3][8] = {
      {0.16009398, -0.094634809, 0.025146376, -0.0030648343, 1, 0.32668115, 0.10406089,
       0.0014179084},
      {0.53487893, 0.1298071, 0.065735949, -0.0015649758, 0.16639465, 0.020070113, -0.0035938915,
       -0.00058392623},
      {0.040121004, -0.0065914022, -0.0026286047, -0.0013441777, 0.017050642, -0.0021309326,
       0.00085092367, -1.5247877e-07},
  };
  double cv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) cv[i] = K[0][i] + u * (K[1][i] + u * K[2][i]);
  const double p = cv[0] + v * (cv[1] + v * (cv[2] + v * cv[3]));
  const double q = cv[4] + v * (cv[5] + v * (cv[6] + v * cv[7]));
  return exp(p / q);
}

// Beta(x; a, b) gradient w.r.t. a for x near 0 (Taylor).        ATen Distributions.h:374-390
__device__ inline double beta_grad_a_small(double x, double a, double b) {
  const double fac = digamma_d(a) - digamma_d(a + b) - log(x);
  double num = 1.0;
  double ser = num / a * (fac + 1.0 / a);
  for (int i = 1; i <= 10; ++i) {
    num *= ((double)i - b) * x / (double)i;
    const double den = a + (double)i;
    ser += num / den * (fac + 1.0 / den);
  }
  const double r = x * pow(1.0 - x, -b) * ser;
  return r != r ? 0.0 : r;
}

// Beta(x; a, b) gradient w.r.t. b for x near 0 (Taylor).        ATen Distributions.h:394-408
__device__ inline double beta_grad_b_small(double x, double a, double b) {
  const double fac = digamma_d(a + b) - digamma_d(b);
  double num = 1.0, bs = 1.0, dbs = 0.0, ser = fac / a;
  for (int i = 1; i <= 8; ++i) {
    num *= -x / (double)i;
    dbs = dbs * (b - (double)i) + bs;
    bs = bs * (b - (double)i);
    ser += num / (a + (double)i) * (dbs + fac * bs);
  }
  const double r = -pow(1.0 - x, 1.0 - b) * ser;
  return r != r ? 0.0 : r;
}

// Both shapes large: Rice saddle-point expansion.                ATen Distributions.h:413-446
__device__ inline double beta_grad_a_mid(double x, double a, double b) {
  const double tot = a + b;
  const double mean = a / tot;
  const double sd = sqrt(a * b / (tot + 1.0)) / tot;
  if (mean - 0.1 * sd <= x && x <= mean + 0.1 * sd) {
    const double b2 = b * b;
    const double poly =
        47.0 * x * b2 * b2 +
        a * ((43.0 + 20.0 * (16.0 + 27.0 * b) * x) * b2 * b +
             a * (3.0 * (59.0 + 180.0 * b - 90.0 * x) * b2 +
                  a * ((453.0 + 1620.0 * b * (1.0 - x) - 455.0 * x) * b +
                       a * (8.0 * (1.0 - x) * (135.0 * b - 11.0)))));
    const double pn = (1.0 + 12.0 * a) * (1.0 + 12.0 * b) / (tot * tot);
    const double pd = 12960.0 * a * a * a * b2 * (1.0 + 12.0 * tot);
    return pn / (1.0 - x) * poly / pd;
  }
  const double pre = -x / sqrt(2.0 * a * b / tot);
  const double stir = (1.0 + 1.0 / (12.0 * a) + 1.0 / (288.0 * a * a)) *
                      (1.0 + 1.0 / (12.0 * b) + 1.0 / (288.0 * b * b)) /
                      (1.0 + 1.0 / (12.0 * tot) + 1.0 / (288.0 * tot * tot));
  const double t1n = 2.0 * (a * a) * (x - 1.0) + a * b * (x - 1.0) - x * (b * b);
  const double axbx = a * (x - 1.0) + b * x;
  const double t1d = sqrt(2.0 * a / b) * pow(tot, 1.5) * axbx * axbx;
  const double t1 = t1n / t1d;
  const double t2 = 0.5 * log(a / (tot * x));
  const double t3 = sqrt(8.0 * a * b / tot) / (b * x + a * (x - 1.0));
  const double t4b = b * log(b / (tot * (1.0 - x))) + a * log(a / (tot * x));
  const double t4 = pow(t4b, -1.5);
  const double t1234 = t1 + t2 * (t3 + (x < mean ? t4 : -t4));
  return stir * pre * t1234;
}

// Scaled gradient  -(dCDF/dalpha)/pdf/(1-x)  of a Beta(alpha, total-alpha) draw x, the quantity
// `torch._dirichlet_grad` returns element-wise.                  ATen Distributions.h:453-511
static __device__ __noinline__ double dirichlet_grad(double x, double alpha, double total) {
  const double beta = total - alpha;
  const double boundary = total * x * (1.0 - x);
  if (x <= 0.5 && boundary < 2.5) return beta_grad_a_small(x, alpha, beta);
  if (x >= 0.5 && boundary < 0.75) return -beta_grad_b_small(1.0 - x, beta, alpha);
  if (alpha > 6.0 && beta > 6.0) return beta_grad_a_mid(x, alpha, beta);

  // Rational correction to an analytic approximation (fitted coefficients).
  const double C[2][3][3][4] = {
      {{{1.003668233, -0.01061107488, -0.0657888334, 0.01201642863},
        {0.6336835991, -0.3557432599, 0.05486251648, -0.001465281033},
        {-0.03276231906, 0.004474107445, 0.002429354597, -0.0001557569013}},
       {{0.221950385, -0.3187676331, 0.01799915743, 0.01074823814},
        {-0.2951249643, 0.06219954479, 0.01535556598, 0.001550077057},
        {0.02155310298, 0.004170831599, 0.001292462449, 6.976601077e-05}},
       {{-0.05980841433, 0.008441916499, 0.01085618172, 0.002319392565},
        {0.02911413504, 0.01400243777, -0.002721828457, 0.000751041181},
        {0.005900514878, -0.001936558688, -9.495446725e-06, 5.385558597e-05}}},
      {{{1, -0.02924021934, -0.04438342661, 0.007285809825},
        {0.6357567472, -0.3473456711, 0.05454656494, -0.002407477521},
        {-0.03301322327, 0.004845219414, 0.00231480583, -0.0002307248149}},
       {{0.5925320577, -0.1757678135, 0.01505928619, 0.000564515273},
        {0.1014815858, -0.06589186703, 0.01272886114, -0.0007316646956},
        {-0.007258481865, 0.001096195486, 0.0003934994223, -4.12701925e-05}},
       {{0.06469649321, -0.0236701437, 0.002902096474, -5.896963079e-05},
        {0.001925008108, -0.002869809258, 0.0008000589141, -6.063713228e-05},
        {-0.0003477407336, 6.959756487e-05, 1.097287507e-05, -1.650964693e-06}}},
  };
  const double u = log(x);
  const double a = log(alpha) - u;
  const double b = log(total) - a;
  const double pu[3] = {1.0, u, u * u};
  const double pa[3] = {1.0, a, a * a};
  double p = 0.0, q = 0.0;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      const double ua = pu[i] * pa[j];
      p += ua * (C[0][i][j][0] + b * (C[0][i][j][1] + b * (C[0][i][j][2] + b * C[0][i][j][3])));
      q += ua * (C[1][i][j][0] + b * (C[1][i][j][1] + b * (C[1][i][j][2] + b * C[1][i][j][3])));
    }
  }
  const double approx = x * (digamma_d(total) - digamma_d(alpha)) / beta;
  return p / q * approx;
}

}  // namespace mnf

/*
 * ql_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * A plain-C restatement of the reference's algorithm for the SCF hot path of
 * QuantumLab.jl (the "stub" pure-Julia path, src/QuantumLab.jl:7-18).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product (libqlb200.so) never links or calls it.
 *
 * Parity status: PINNED for s/p shells by the reference's own known-answer tests
 * (test/runtests.jl:70,84,93,120,148; see tests/test_oracle_kat.py).  For d shells the
 * reference has no test and its Boys routine is numerically broken, so the
 * oracle of record there is this restated closed form evaluated with an exact Boys
 * function ("parity unpinned by the reference" for d-shell configurations).
 *
 * Two integrators live here:
 *   1. cook_*  : literal restatement of IntegralsModule.jl:66-120 (Gaussian product,
 *                polynomial factors), :283-289 (FIntegral), :354-409 (theta/B factors),
 *                :412-462 (primitive / contracted ERI).  Oracle of record.
 *   2. md_*    : McMurchie-Davidson integrator with 8-fold symmetry + integer Schwarz x
 *                density screening, OpenMP.  Validated against (1) by tests; used as the
 *                oracle for large configurations and as the "fair CPU" baseline.
 *
 * Conventions follow the reference: Cartesian components ordered x descending then y
 * descending (BaseModule.jl:110-122); shell coefficients normalised as in the Shell
 * constructor (ShellModule.jl:25-48); non-axial factors as in ShellModule.jl:50-58;
 * all matrices/tensors column-major, first index fastest.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXL 4            /* per-shell L supported by the oracle (s..g)            */
#define MAXLT (4 * MAXL)  /* total L of a quartet                                   */
#define NCART(l) (((l) + 1) * ((l) + 2) / 2)

/* ------------------------------------------------------------------ helpers */
static double dfact(int n) { /* BaseModule.jl:130  doublefactorial(n)=prod(n:-2:1) */
  double r = 1.0;
  for (int k = n; k >= 1; k -= 2) r *= k;
  return r;
}
static double fact(int n) {
  double r = 1.0;
  for (int k = 2; k <= n; k++) r *= k;
  return r;
}
static double binom(int n, int k) {
  if (k < 0 || k > n) return 0.0;
  return fact(n) / (fact(k) * fact(n - k));
}
static double ipow(double x, int n) { /* Julia x^n with 0.0^0 == 1.0 */
  double r = 1.0;
  for (int k = 0; k < n; k++) r *= x;
  return r;
}
/* Cartesian component list of a shell, BaseModule.jl:110-122 */
static int cart_list(int L, int out[][3]) {
  int i = 0;
  for (int x = L; x >= 0; x--)
    for (int y = L - x; y >= 0; y--) {
      out[i][0] = x; out[i][1] = y; out[i][2] = L - x - y; i++;
    }
  return i;
}
/* ShellModule.jl:50-58 computeFactorsNonAxial */
static double nonaxial(int L, const int m[3]) {
  double na = dfact(2 * m[0] - 1) * dfact(2 * m[1] - 1) * dfact(2 * m[2] - 1) / dfact(2 * L - 1);
  return 1.0 / sqrt(na);
}

/* ------------------------------------------------------------------ Boys function */
/* exact: series for small x, erf + upward recursion for large x */
static void boys_exact_array(int mmax, double x, double *F) {
  if (x < 35.0) {
    /* F_m(x) = exp(-x) sum_k (2x)^k / ((2m+1)(2m+3)...(2m+2k+1)) for m = mmax, then downward */
    double term = 1.0 / (2 * mmax + 1), sum = term;
    for (int k = 1; k < 400; k++) {
      term *= 2.0 * x / (2 * mmax + 2 * k + 1);
      sum += term;
      if (term < 1e-17 * sum) break;
    }
    double ex = exp(-x);
    F[mmax] = ex * sum;
    for (int m = mmax; m > 0; m--) F[m - 1] = (2.0 * x * F[m] + ex) / (2 * m - 1);
  } else {
    double ex = exp(-x);
    F[0] = 0.5 * sqrt(M_PI / x) * erf(sqrt(x));
    for (int m = 0; m < mmax; m++) F[m + 1] = ((2 * m + 1) * F[m] - ex) / (2.0 * x);
  }
}
double orc_boys_exact(int m, double x) {
  double F[64];
  boys_exact_array(m, x, F);
  return F[m];
}
/* upper incomplete gamma for half-integer a, the quantity GSL.sf_gamma_inc returns */
static double gamma_inc_upper(double a, double x) {
  if (x < a + 1.0) { /* lower series, then Gamma(a) - lower */
    double term = 1.0 / a, sum = term;
    for (int k = 1; k < 1000; k++) {
      term *= x / (a + k);
      sum += term;
      if (term < 1e-17 * sum) break;
    }
    double lower = sum * exp(-x + a * log(x));
    return tgamma(a) - lower;
  } else { /* Lentz continued fraction */
    double b = x + 1.0 - a, c = 1.0 / 1e-300, d = 1.0 / b, h = d;
    for (int i = 1; i < 1000; i++) {
      double an = -i * (i - a);
      b += 2.0;
      d = an * d + b; if (fabs(d) < 1e-300) d = 1e-300;
      c = b + an / c; if (fabs(c) < 1e-300) c = 1e-300;
      d = 1.0 / d;
      double del = d * c;
      h *= del;
      if (fabs(del - 1.0) < 1e-16) break;
    }
    return exp(-x + a * log(x)) * h;
  }
}
/* IntegralsModule.jl:283-289 FIntegral, literal (incl. the clamp and the Gamma difference) */
double orc_boys_refquirk(int m, double x) {
  const double APPROX_ZERO = 1e-5;
  if (x < APPROX_ZERO) x = APPROX_ZERO;
  if (x <= APPROX_ZERO) return 1.0 / (1 + 2 * m);
  return 0.5 * pow(x, -0.5 - m) * (tgamma(0.5 + m) - gamma_inc_upper(0.5 + m, x));
}
static double boys(int m, double x, int mode) {
  return mode == 1 ? orc_boys_refquirk(m, x) : orc_boys_exact(m, x);
}

/* ------------------------------------------------------------------ primitives */
typedef struct {
  double c[3];  /* centre */
  double a;     /* exponent */
  int l[3];     /* Cartesian powers (MQuantumNumber) */
} pgbf;

/* IntegralsModule.jl:66-83 GaussProductFundamental */
static void gauss_product(const pgbf *p1, const pgbf *p2, double *K, double P[3], double *g) {
  double a1 = p1->a, a2 = p2->a;
  double gamma = a1 + a2, ab2 = 0.0;
  for (int k = 0; k < 3; k++) {
    P[k] = (a1 * p1->c[k] + a2 * p2->c[k]) / gamma;
    double d = p1->c[k] - p2->c[k];
    ab2 += d * d;
  }
  /* reference: AB = distance(A,B) (a sqrt), then AB^2 */
  double AB = sqrt(ab2);
  *K = exp(-a1 * a2 * AB * AB / gamma);
  *g = gamma;
}
/* IntegralsModule.jl:91-120 GaussProductPolynomialFactor, one axis: f[k], k=0..l1+l2 */
static void poly_factor(int l1, int l2, double PA, double PB, double *f) {
  for (int k = 0; k <= l1 + l2; k++) {
    double s = 0.0;
    int qlo = (-k > k - 2 * l2) ? -k : k - 2 * l2;
    int qhi = (k < 2 * l1 - k) ? k : 2 * l1 - k;
    for (int q = qlo; q <= qhi; q += 2) {
      int i = (k + q) / 2, j = (k - q) / 2;
      s += binom(l1, i) * binom(l2, j) * ipow(PA, l1 - i) * ipow(PB, l2 - j);
    }
    f[k] = s;
  }
}

/* theta list for one axis: IntegralsModule.jl:360-378 */
typedef struct { double th; int l, r; } theta_t;
static int theta_axis(int l1, int l2, double PA, double PB, double gamma, theta_t *out) {
  double f[2 * MAXL + 1];
  poly_factor(l1, l2, PA, PB, f);
  int n = 0;
  for (int l = 0; l <= l1 + l2; l++)
    for (int r = 0; r <= l / 2; r++) {
      out[n].th = f[l] * (fact(l) * pow(gamma, (double)(r - l))) / (fact(r) * fact(l - 2 * r));
      out[n].l = l; out[n].r = r; n++;
    }
  return n;
}
#define MAXTH 32
#define MAXB 1024
typedef struct { double B; int v; } bterm_t; /* v = l12+l34-2(r12+r34)-i : Boys-order share */

/* IntegralsModule.jl:386-409 BFactors, one axis */
static int b_axis(const theta_t *t12, int n12, const theta_t *t34, int n34, double delta, double p, bterm_t *out) {
  int n = 0;
  for (int a = 0; a < n12; a++)
    for (int b = 0; b < n34; b++) {
      int l12 = t12[a].l, r12 = t12[a].r, l34 = t34[b].l, r34 = t34[b].r;
      int top = l12 + l34 - 2 * r12 - 2 * r34;
      for (int i = 0; i <= top / 2; i++) {
        double sgn = ((l34 & 1) ? -1.0 : 1.0) * ((i & 1) ? -1.0 : 1.0);
        double B = sgn * t12[a].th * t34[b].th *
                   (pow(2 * delta, 2.0 * (r12 + r34)) * fact(top) * ipow(delta, i) * ipow(p, top - 2 * i)) /
                   (pow(4 * delta, (double)(l12 + l34)) * fact(i) * fact(top - 2 * i));
        out[n].B = B; out[n].v = top - i; n++;
      }
    }
  return n;
}

/* IntegralsModule.jl:412-445 primitive ERI (Mulliken order), Cook eq. 7.1.
 * literal != 0 : triple loop over (Bx,By,Bz) with one Boys call per term, as the reference does.
 * literal == 0 : same closed form with the per-axis terms grouped by Boys order (re-association only). */
static double cook_prim(const pgbf *mu, const pgbf *nu, const pgbf *la, const pgbf *si, int boys_mode, int literal) {
  double K1, K2, P[3], Q[3], g1, g2;
  gauss_product(mu, nu, &K1, P, &g1);
  gauss_product(la, si, &K2, Q, &g2);
  double delta = 1.0 / (4 * g1) + 1.0 / (4 * g2);
  double pq2 = 0.0;
  for (int k = 0; k < 3; k++) pq2 += (P[k] - Q[k]) * (P[k] - Q[k]);
  double PQ = sqrt(pq2);
  double x = PQ * PQ / (4 * delta);
  double ab2 = 0, cd2 = 0;
  for (int k = 0; k < 3; k++) {
    ab2 += (mu->c[k] - nu->c[k]) * (mu->c[k] - nu->c[k]);
    cd2 += (la->c[k] - si->c[k]) * (la->c[k] - si->c[k]);
  }
  double AB = sqrt(ab2), CD = sqrt(cd2);
  double Omega = 2 * M_PI * M_PI / (g1 * g2) * sqrt(M_PI / (g1 + g2)) *
                 exp(-mu->a * nu->a * AB * AB / g1 - la->a * si->a * CD * CD / g2);
  static __thread bterm_t Bt[3][MAXB];
  int nB[3];
  for (int ax = 0; ax < 3; ax++) {
    theta_t t12[MAXTH], t34[MAXTH];
    int n12 = theta_axis(mu->l[ax], nu->l[ax], P[ax] - mu->c[ax], P[ax] - nu->c[ax], g1, t12);
    int n34 = theta_axis(la->l[ax], si->l[ax], Q[ax] - la->c[ax], Q[ax] - si->c[ax], g2, t34);
    nB[ax] = b_axis(t12, n12, t34, n34, delta, Q[ax] - P[ax], Bt[ax]);
  }
  double result = 0.0;
  if (literal) {
    for (int i = 0; i < nB[0]; i++)
      for (int j = 0; j < nB[1]; j++)
        for (int k = 0; k < nB[2]; k++) {
          int V = Bt[0][i].v + Bt[1][j].v + Bt[2][k].v;
          result += Omega * Bt[0][i].B * Bt[1][j].B * Bt[2][k].B * boys(V, x, boys_mode);
        }
  } else {
    double S[3][MAXLT + 1];
    int vmax[3] = {0, 0, 0};
    memset(S, 0, sizeof S);
    for (int ax = 0; ax < 3; ax++)
      for (int i = 0; i < nB[ax]; i++) {
        S[ax][Bt[ax][i].v] += Bt[ax][i].B;
        if (Bt[ax][i].v > vmax[ax]) vmax[ax] = Bt[ax][i].v;
      }
    double F[3 * MAXLT + 1];
    int VM = vmax[0] + vmax[1] + vmax[2];
    if (boys_mode == 1) for (int m = 0; m <= VM; m++) F[m] = orc_boys_refquirk(m, x);
    else boys_exact_array(VM, x, F);
    for (int i = 0; i <= vmax[0]; i++)
      for (int j = 0; j <= vmax[1]; j++)
        for (int k = 0; k <= vmax[2]; k++) result += Omega * S[0][i] * S[1][j] * S[2][k] * F[i + j + k];
  }
  return result;
}

/* ------------------------------------------------------------------ basis plumbing */
typedef struct {
  int nsh;
  const double *centers; /* 3*nsh */
  const int *L;          /* nsh */
  const int *nprim;      /* nsh */
  const double *exps;    /* concatenated */
  const double *coefs;   /* concatenated, already normalised per ShellModule.jl:25-48 */
  int *poff;             /* nsh+1 primitive offsets */
  int *boff;             /* nsh+1 basis-function offsets */
} basis_t;

static void basis_init(basis_t *b, int nsh, const double *centers, const int *L, const int *nprim, const double *exps,
                       const double *coefs) {
  b->nsh = nsh; b->centers = centers; b->L = L; b->nprim = nprim; b->exps = exps; b->coefs = coefs;
  b->poff = (int *)malloc(sizeof(int) * (nsh + 1));
  b->boff = (int *)malloc(sizeof(int) * (nsh + 1));
  b->poff[0] = 0; b->boff[0] = 0;
  for (int i = 0; i < nsh; i++) {
    b->poff[i + 1] = b->poff[i] + nprim[i];
    b->boff[i + 1] = b->boff[i] + NCART(L[i]);
  }
}
static void basis_free(basis_t *b) { free(b->poff); free(b->boff); }

/* ShellModule.jl:25-48: Shell constructor renormalisation (in place, like the reference) */
void orc_shell_renorm(int L, int nprim, const double *exps, double *coefs) {
  for (int p = 0; p < nprim; p++)
    coefs[p] *= sqrt((pow(2.0, L) * pow(2 * exps[p], L + 1.5)) / (pow(M_PI, 1.5) * dfact(2 * L - 1)));
  double norm = 0.0;
  for (int p = 0; p < nprim; p++)
    for (int q = 0; q < nprim; q++)
      norm += (dfact(2 * L - 1) * pow(M_PI, 1.5) * coefs[p] * coefs[q]) / (pow(2.0, L) * pow(exps[p] + exps[q], L + 1.5));
  for (int p = 0; p < nprim; p++) coefs[p] /= sqrt(norm);
}

/* contracted Cook ERI between component ia of shell A ... (IntegralsModule.jl:447-462 with the
 * expandShell coefficients of ShellModule.jl:119-128 = shell coefficient * non-axial factor) */
static double cook_contracted(const basis_t *b, const int sh[4], int comp[4][3], int boys_mode, int literal) {
  pgbf g[4];
  double na = 1.0;
  for (int s = 0; s < 4; s++) {
    for (int k = 0; k < 3; k++) { g[s].c[k] = b->centers[3 * sh[s] + k]; g[s].l[k] = comp[s][k]; }
    na *= nonaxial(b->L[sh[s]], comp[s]);
  }
  double sum = 0.0;
  for (int p0 = 0; p0 < b->nprim[sh[0]]; p0++)
    for (int p1 = 0; p1 < b->nprim[sh[1]]; p1++)
      for (int p2 = 0; p2 < b->nprim[sh[2]]; p2++)
        for (int p3 = 0; p3 < b->nprim[sh[3]]; p3++) {
          int pi[4] = {p0, p1, p2, p3};
          double c = 1.0;
          for (int s = 0; s < 4; s++) {
            g[s].a = b->exps[b->poff[sh[s]] + pi[s]];
            c *= b->coefs[b->poff[sh[s]] + pi[s]];
          }
          sum += c * cook_prim(&g[0], &g[1], &g[2], &g[3], boys_mode, literal);
        }
  return sum * na;
}

/* SpecialMatricesModule.jl:184-186 computeTensorBlockElectronRepulsionIntegrals(Shell x4):
 * out[mu + nA*(nu + nB*(la + nC*si))] */
static void cook_block(const basis_t *b, int i, int j, int k, int l, int boys_mode, int literal, double *out) {
  int ca[15][3], cb[15][3], cc[15][3], cd[15][3];
  int nA = cart_list(b->L[i], ca), nB = cart_list(b->L[j], cb), nC = cart_list(b->L[k], cc), nD = cart_list(b->L[l], cd);
  int sh[4] = {i, j, k, l};
  for (int d = 0; d < nD; d++)
    for (int c = 0; c < nC; c++)
      for (int bb = 0; bb < nB; bb++)
        for (int a = 0; a < nA; a++) {
          int comp[4][3];
          memcpy(comp[0], ca[a], sizeof(int) * 3); memcpy(comp[1], cb[bb], sizeof(int) * 3);
          memcpy(comp[2], cc[c], sizeof(int) * 3); memcpy(comp[3], cd[d], sizeof(int) * 3);
          out[a + nA * (bb + nB * (c + nC * d))] = cook_contracted(b, sh, comp, boys_mode, literal);
        }
}
void orc_cook_eri_block(int nsh, const double *centers, const int *L, const int *nprim, const double *exps,
                        const double *coefs, int i, int j, int k, int l, int boys_mode, int literal, double *out) {
  basis_t b; basis_init(&b, nsh, centers, L, nprim, exps, coefs);
  cook_block(&b, i, j, k, l, boys_mode, literal, out);
  basis_free(&b);
}
/* a bare primitive ERI for formula-level tests */
double orc_cook_eri_prim(const double *centers /*4x3*/, const double *alphas /*4*/, const int *lmn /*4x3*/, int boys_mode,
                         int literal) {
  pgbf g[4];
  for (int s = 0; s < 4; s++) {
    for (int k = 0; k < 3; k++) { g[s].c[k] = centers[3 * s + k]; g[s].l[k] = lmn[3 * s + k]; }
    g[s].a = alphas[s];
  }
  return cook_prim(&g[0], &g[1], &g[2], &g[3], boys_mode, literal);
}

/* SpecialMatricesModule.jl:198-207 full tensor from shells, out[mu,nu,la,si] column-major N^4 */
void orc_cook_eri_tensor(int nsh, const double *centers, const int *L, const int *nprim, const double *exps,
                         const double *coefs, int boys_mode, int literal, double *out) {
  basis_t b; basis_init(&b, nsh, centers, L, nprim, exps, coefs);
  long N = b.boff[nsh];
#pragma omp parallel for collapse(2) schedule(dynamic)
  for (int i = 0; i < nsh; i++)
    for (int j = 0; j < nsh; j++) {
      double blk[1296];
      for (int k = 0; k < nsh; k++)
        for (int l = 0; l < nsh; l++) {
          int nA = NCART(L[i]), nB = NCART(L[j]), nC = NCART(L[k]), nD = NCART(L[l]);
          if ((long)nA * nB * nC * nD > 1296) continue; /* oracle tensor limited to s,p,d */
          cook_block(&b, i, j, k, l, boys_mode, literal, blk);
          for (int d = 0; d < nD; d++)
            for (int c = 0; c < nC; c++)
              for (int bb = 0; bb < nB; bb++)
                for (int a = 0; a < nA; a++)
                  out[(b.boff[i] + a) + N * ((b.boff[j] + bb) + N * ((b.boff[k] + c) + N * (long)(b.boff[l] + d)))] =
                      blk[a + nA * (bb + nB * (c + nC * d))];
        }
    }
  basis_free(&b);
}

/* SpecialMatricesModule.jl:33-69 contractTensorBlocksWithMatrix_CoulombLike and :71-100 _ExchangeLike:
 * all nsh^4 ORDERED quartets, no symmetry, no screening, one pass for J and a separate pass for K.
 * which: 0 = Coulomb, 1 = Exchange.  quartet_stride>1 samples every stride-th quartet (bench sampling);
 * *nquartets returns the number of shell quartets actually evaluated. */
void orc_ref_contract(int nsh, const double *centers, const int *L, const int *nprim, const double *exps,
                      const double *coefs, const double *P, int which, int boys_mode, int literal, long quartet_stride,
                      double *out, long *nquartets) {
  basis_t b; basis_init(&b, nsh, centers, L, nprim, exps, coefs);
  long N = b.boff[nsh];
  memset(out, 0, sizeof(double) * N * N);
  long count = 0, seen = 0;
  double blk[1296];
  for (int s1 = 0; s1 < nsh; s1++)
    for (int s2 = 0; s2 < nsh; s2++)
      for (int s3 = 0; s3 < nsh; s3++)
        for (int s4 = 0; s4 < nsh; s4++) {
          if (quartet_stride > 1 && (seen++ % quartet_stride) != 0) continue;
          int n1 = NCART(L[s1]), n2 = NCART(L[s2]), n3 = NCART(L[s3]), n4 = NCART(L[s4]);
          count++;
          if (which == 0) {
            cook_block(&b, s1, s2, s3, s4, boys_mode, literal, blk); /* tensor[mu,nu,la,si] */
            for (int mu = 0; mu < n1; mu++)
              for (int nu = 0; nu < n2; nu++) {
                double acc = 0.0;
                for (int si = 0; si < n4; si++)
                  for (int la = 0; la < n3; la++)
                    acc += blk[mu + n1 * (nu + n2 * (la + n3 * si))] * P[(b.boff[s3] + la) + N * (b.boff[s4] + si)];
                out[(b.boff[s1] + mu) + N * (b.boff[s2] + nu)] += acc;
              }
          } else {
            cook_block(&b, s1, s3, s2, s4, boys_mode, literal, blk); /* tensor[mu,la,nu,si] */
            for (int mu = 0; mu < n1; mu++)
              for (int nu = 0; nu < n2; nu++) {
                double acc = 0.0;
                for (int si = 0; si < n4; si++)
                  for (int la = 0; la < n3; la++)
                    acc += blk[mu + n1 * (la + n3 * (nu + n2 * si))] * P[(b.boff[s3] + la) + N * (b.boff[s4] + si)];
                out[(b.boff[s1] + mu) + N * (b.boff[s2] + nu)] += acc;
              }
          }
        }
  if (nquartets) *nquartets = count;
  basis_free(&b);
}

/* Sampled, optionally multi-threaded timing form of the reference's direct contraction
 * (SpecialMatricesModule.jl:33-100): evaluates every `stride`-th quartet of the ordered nsh^4 loop (uniform sample),
 * for J (which=0) or K (which=1), with `nthreads` OpenMP threads (the reference itself is single-threaded; >1 models
 * the reference algorithm spread over all host cores).  Returns the number of shell quartets evaluated; `out` gets
 * the (partial) contraction so the work cannot be optimised away. */
long orc_ref_contract_sampled(int nsh, const double *centers, const int *L, const int *nprim, const double *exps,
                              const double *coefs, const double *P, int which, int boys_mode, int literal, long stride,
                              long offset, int nthreads, double *out) {
  basis_t b; basis_init(&b, nsh, centers, L, nprim, exps, coefs);
  long N = b.boff[nsh];
  memset(out, 0, sizeof(double) * N * N);
  long total = (long)nsh * nsh * nsh * nsh;
  long nsample = (total - offset + stride - 1) / stride;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    double *loc = (double *)calloc(N * N, sizeof(double));
    double blk[1296];
#pragma omp for schedule(dynamic, 16)
    for (long s = 0; s < nsample; s++) {
      long q = offset + s * stride;
      int s4 = (int)(q % nsh); q /= nsh;
      int s3 = (int)(q % nsh); q /= nsh;
      int s2 = (int)(q % nsh); q /= nsh;
      int s1 = (int)q;
      int n1 = NCART(L[s1]), n2 = NCART(L[s2]), n3 = NCART(L[s3]), n4 = NCART(L[s4]);
      if (which == 0) {
        cook_block(&b, s1, s2, s3, s4, boys_mode, literal, blk);
        for (int mu = 0; mu < n1; mu++)
          for (int nu = 0; nu < n2; nu++) {
            double acc = 0.0;
            for (int si = 0; si < n4; si++)
              for (int la = 0; la < n3; la++)
                acc += blk[mu + n1 * (nu + n2 * (la + n3 * si))] * P[(b.boff[s3] + la) + N * (b.boff[s4] + si)];
            loc[(b.boff[s1] + mu) + N * (b.boff[s2] + nu)] += acc;
          }
      } else {
        cook_block(&b, s1, s3, s2, s4, boys_mode, literal, blk);
        for (int mu = 0; mu < n1; mu++)
          for (int nu = 0; nu < n2; nu++) {
            double acc = 0.0;
            for (int si = 0; si < n4; si++)
              for (int la = 0; la < n3; la++)
                acc += blk[mu + n1 * (la + n3 * (nu + n2 * si))] * P[(b.boff[s3] + la) + N * (b.boff[s4] + si)];
            loc[(b.boff[s1] + mu) + N * (b.boff[s2] + nu)] += acc;
          }
      }
    }
#pragma omp critical
    for (long t = 0; t < N * N; t++) out[t] += loc[t];
    free(loc);
  }
  basis_free(&b);
  return nsample;
}

/* ------------------------------------------------------------------ one-electron integrals (for the KATs) */
/* IntegralsModule.jl:39-49 GaussianIntegral1D_Mathematica */
static double gint1d(int m, double z) {
  if (m % 2 != 0) return 0.0;
  double t = (m + 1) / 2.0;
  return pow(z, -t) * tgamma(t);
}
/* IntegralsModule.jl:138-150 */
static double overlap_prim(const pgbf *p1, const pgbf *p2) {
  double K, P[3], g, I[3];
  if (p1->l[0] < 0 || p1->l[1] < 0 || p1->l[2] < 0 || p2->l[0] < 0 || p2->l[1] < 0 || p2->l[2] < 0) return 0.0;
  gauss_product(p1, p2, &K, P, &g);
  for (int ax = 0; ax < 3; ax++) {
    double f[2 * MAXL + 3];
    poly_factor(p1->l[ax], p2->l[ax], P[ax] - p1->c[ax], P[ax] - p2->c[ax], f);
    double s = 0.0;
    for (int k = 0; k <= p1->l[ax] + p2->l[ax]; k++) s += f[k] * gint1d(k, g);
    I[ax] = s;
  }
  return K * I[0] * I[1] * I[2];
}
/* IntegralsModule.jl:228-260 */
static double kinetic_prim(const pgbf *p1, const pgbf *p2) {
  double integral = 0.0;
  for (int ax = 0; ax < 3; ax++) {
    pgbf d1 = *p1, i1 = *p1, d2 = *p2, i2 = *p2;
    d1.l[ax] -= 1; i1.l[ax] += 1; d2.l[ax] -= 1; i2.l[ax] += 1;
    int l1 = p1->l[ax], l2 = p2->l[ax];
    double a1 = p1->a, a2 = p2->a;
    if (l1 * l2 != 0) integral += (0.5 * l1 * l2) * overlap_prim(&d1, &d2);
    integral += (2 * a1 * a2) * overlap_prim(&i1, &i2);
    if (l2 != 0) integral += (-a1 * l2) * overlap_prim(&i1, &d2);
    if (l1 != 0) integral += (-l1 * a2) * overlap_prim(&d1, &i2);
  }
  return integral;
}
/* IntegralsModule.jl:291-304 ASummand (the f factor carries the sign convention of the reference) */
static double asummand(double f, double PC, double g, int l, int r, int i) {
  double sgn = ((l & 1) ? -1.0 : 1.0) * ((i & 1) ? -1.0 : 1.0);
  return sgn * f * fact(l) * ipow(PC, l - 2 * r - 2 * i) * ipow(1.0 / (4 * g), r + i) / (fact(r) * fact(i) * fact(l - 2 * r - 2 * i));
}
/* IntegralsModule.jl:306-338 */
static double nuclear_prim(const pgbf *p1, const pgbf *p2, const double C[3], double Z, int boys_mode) {
  double K, P[3], g;
  gauss_product(p1, p2, &K, P, &g);
  double pc2 = 0.0;
  for (int k = 0; k < 3; k++) pc2 += (P[k] - C[k]) * (P[k] - C[k]);
  double PCd = sqrt(pc2);
  double f[3][2 * MAXL + 3];
  for (int ax = 0; ax < 3; ax++) poly_factor(p1->l[ax], p2->l[ax], P[ax] - p1->c[ax], P[ax] - p2->c[ax], f[ax]);
  double result = 0.0;
  int Lx = p1->l[0] + p2->l[0], Ly = p1->l[1] + p2->l[1], Lz = p1->l[2] + p2->l[2];
  for (int l = 0; l <= Lx; l++) for (int r = 0; r <= l / 2; r++) for (int i = 0; i <= (l - 2 * r) / 2; i++) {
    double Ax = asummand(f[0][l], P[0] - C[0], g, l, r, i);
    for (int m = 0; m <= Ly; m++) for (int s = 0; s <= m / 2; s++) for (int j = 0; j <= (m - 2 * s) / 2; j++) {
      double Ay = asummand(f[1][m], P[1] - C[1], g, m, s, j);
      for (int n = 0; n <= Lz; n++) for (int t = 0; t <= n / 2; t++) for (int k = 0; k <= (n - 2 * t) / 2; k++) {
        double Az = asummand(f[2][n], P[2] - C[2], g, n, t, k);
        result += 2 * M_PI / g * K * Ax * Ay * Az * boys(l + m + n - 2 * (r + s + t) - (i + j + k), g * PCd * PCd, boys_mode);
      }
    }
  }
  return -Z * result;
}

/* which: 0 overlap, 1 kinetic, 2 nuclear attraction (atoms: natoms x {x,y,z}, Z) ; out N x N.
 * SpecialMatricesModule.jl:102-177 (shell form via scatterMatrixBlocks2D) */
void orc_one_electron(int nsh, const double *centers, const int *L, const int *nprim, const double *exps,
                      const double *coefs, int which, int natoms, const double *atom_xyz, const double *atom_Z,
                      int boys_mode, double *out) {
  basis_t b; basis_init(&b, nsh, centers, L, nprim, exps, coefs);
  long N = b.boff[nsh];
#pragma omp parallel for schedule(dynamic)
  for (int i = 0; i < nsh; i++)
    for (int j = 0; j < nsh; j++) {
      int ca[15][3], cb[15][3];
      int nA = cart_list(L[i], ca), nB = cart_list(L[j], cb);
      for (int a = 0; a < nA; a++)
        for (int bb = 0; bb < nB; bb++) {
          pgbf g1, g2;
          for (int k = 0; k < 3; k++) {
            g1.c[k] = centers[3 * i + k]; g1.l[k] = ca[a][k];
            g2.c[k] = centers[3 * j + k]; g2.l[k] = cb[bb][k];
          }
          double na = nonaxial(L[i], ca[a]) * nonaxial(L[j], cb[bb]);
          double sum = 0.0;
          for (int p = 0; p < nprim[i]; p++)
            for (int q = 0; q < nprim[j]; q++) {
              g1.a = exps[b.poff[i] + p]; g2.a = exps[b.poff[j] + q];
              double c = coefs[b.poff[i] + p] * coefs[b.poff[j] + q];
              double v = 0.0;
              if (which == 0) v = overlap_prim(&g1, &g2);
              else if (which == 1) v = kinetic_prim(&g1, &g2);
              else for (int at = 0; at < natoms; at++) v += nuclear_prim(&g1, &g2, atom_xyz + 3 * at, atom_Z[at], boys_mode);
              sum += c * v;
            }
          out[(b.boff[i] + a) + N * (b.boff[j] + bb)] = sum * na;
        }
    }
  basis_free(&b);
}

/* ================================================================== McMurchie-Davidson integrator */
#define NHERM(l) (((l) + 1) * ((l) + 2) * ((l) + 3) / 6)
typedef struct {
  double p, P[3], cK; /* exponent sum, centre, c_a*c_b*exp(-ab/p |AB|^2) */
  double E[3][MAXL + 1][MAXL + 1][2 * MAXL + 1]; /* E^{ij}_t per axis */
} ppair_t;

static void hermite_E(int la, int lb, double PA, double PB, double oo2p, double E[MAXL + 1][MAXL + 1][2 * MAXL + 1]) {
  memset(E, 0, sizeof(double) * (MAXL + 1) * (MAXL + 1) * (2 * MAXL + 1));
  E[0][0][0] = 1.0;
  for (int i = 0; i <= la; i++)
    for (int j = 0; j <= lb; j++) {
      if (i == 0 && j == 0) continue;
      for (int t = 0; t <= i + j; t++) {
        double v;
        if (i > 0) {
          v = PA * E[i - 1][j][t];
          if (t > 0) v += oo2p * E[i - 1][j][t - 1];
          if (t + 1 <= i - 1 + j) v += (t + 1) * E[i - 1][j][t + 1];
        } else {
          v = PB * E[i][j - 1][t];
          if (t > 0) v += oo2p * E[i][j - 1][t - 1];
          if (t + 1 <= i + j - 1) v += (t + 1) * E[i][j - 1][t + 1];
        }
        E[i][j][t] = v;
      }
    }
}
static void make_ppair(const basis_t *b, int sa, int sb, int pa, int pb, ppair_t *pp) {
  double a = b->exps[b->poff[sa] + pa], bb = b->exps[b->poff[sb] + pb];
  const double *A = b->centers + 3 * sa, *B = b->centers + 3 * sb;
  double p = a + bb, ab2 = 0.0;
  for (int k = 0; k < 3; k++) { pp->P[k] = (a * A[k] + bb * B[k]) / p; ab2 += (A[k] - B[k]) * (A[k] - B[k]); }
  pp->p = p;
  pp->cK = b->coefs[b->poff[sa] + pa] * b->coefs[b->poff[sb] + pb] * exp(-a * bb / p * ab2);
  for (int k = 0; k < 3; k++) hermite_E(b->L[sa], b->L[sb], pp->P[k] - A[k], pp->P[k] - B[k], 0.5 / p, pp->E[k]);
}
/* Hermite Coulomb integrals R_{tuv} for t+u+v<=L, indexed [t][u][v] */
#define RD (MAXLT + 1)
static void hermite_R(int L, double alpha, const double PQ[3], double *R /* RD^3 */) {
  double T = alpha * (PQ[0] * PQ[0] + PQ[1] * PQ[1] + PQ[2] * PQ[2]);
  double F[MAXLT + 1];
  boys_exact_array(L, T, F);
  /* Rn[n][t][u][v] built by downward n */
  static __thread double W[MAXLT + 1][RD][RD][RD];
  double m2a = 1.0;
  for (int n = 0; n <= L; n++) { W[n][0][0][0] = m2a * F[n]; m2a *= -2.0 * alpha; }
  for (int n = L - 1; n >= 0; n--) {
    int lim = L - n;
    for (int t = 0; t <= lim; t++)
      for (int u = 0; u + t <= lim; u++)
        for (int v = 0; v + u + t <= lim; v++) {
          if (t + u + v == 0) continue;
          double val;
          if (t > 0) {
            val = PQ[0] * W[n + 1][t - 1][u][v];
            if (t > 1) val += (t - 1) * W[n + 1][t - 2][u][v];
          } else if (u > 0) {
            val = PQ[1] * W[n + 1][t][u - 1][v];
            if (u > 1) val += (u - 1) * W[n + 1][t][u - 2][v];
          } else {
            val = PQ[2] * W[n + 1][t][u][v - 1];
            if (v > 1) val += (v - 1) * W[n + 1][t][u][v - 2];
          }
          W[n][t][u][v] = val;
        }
  }
  for (int t = 0; t <= L; t++)
    for (int u = 0; u + t <= L; u++)
      for (int v = 0; v + u + t <= L; v++) R[(t * RD + u) * RD + v] = W[0][t][u][v];
}

/* contracted shell-quartet block by MD: out[a + nA*(b + nB*(c + nC*d))] */
static void md_block(const basis_t *b, int sa, int sb, int sc, int sd, double *out) {
  int la = b->L[sa], lb = b->L[sb], lc = b->L[sc], ld = b->L[sd];
  int ca[15][3], cb[15][3], cc[15][3], cd[15][3];
  int nA = cart_list(la, ca), nB = cart_list(lb, cb), nC = cart_list(lc, cc), nD = cart_list(ld, cd);
  int Ltot = la + lb + lc + ld;
  memset(out, 0, sizeof(double) * nA * nB * nC * nD);
  static __thread double R[RD * RD * RD];
  static __thread ppair_t bra, ket;
  for (int pa = 0; pa < b->nprim[sa]; pa++)
    for (int pb = 0; pb < b->nprim[sb]; pb++) {
      make_ppair(b, sa, sb, pa, pb, &bra);
      for (int pc = 0; pc < b->nprim[sc]; pc++)
        for (int pd = 0; pd < b->nprim[sd]; pd++) {
          make_ppair(b, sc, sd, pc, pd, &ket);
          double p = bra.p, q = ket.p, alpha = p * q / (p + q);
          double PQ[3] = {bra.P[0] - ket.P[0], bra.P[1] - ket.P[1], bra.P[2] - ket.P[2]};
          double pref = 2.0 * pow(M_PI, 2.5) / (p * q * sqrt(p + q)) * bra.cK * ket.cK;
          hermite_R(Ltot, alpha, PQ, R);
          for (int d = 0; d < nD; d++)
            for (int c = 0; c < nC; c++) {
              /* G[t][u][v] = sum_q (-1)^q Ecd_q R_{p+q} */
              double G[2 * MAXL + 1][2 * MAXL + 1][2 * MAXL + 1];
              int lab = la + lb;
              for (int t = 0; t <= lab; t++)
                for (int u = 0; u + t <= lab; u++)
                  for (int v = 0; v + u + t <= lab; v++) {
                    double s = 0.0;
                    for (int tt = 0; tt <= cc[c][0] + cd[d][0]; tt++)
                      for (int uu = 0; uu <= cc[c][1] + cd[d][1]; uu++)
                        for (int vv = 0; vv <= cc[c][2] + cd[d][2]; vv++) {
                          double e = ket.E[0][cc[c][0]][cd[d][0]][tt] * ket.E[1][cc[c][1]][cd[d][1]][uu] *
                                     ket.E[2][cc[c][2]][cd[d][2]][vv];
                          if ((tt + uu + vv) & 1) e = -e;
                          s += e * R[((t + tt) * RD + (u + uu)) * RD + (v + vv)];
                        }
                    G[t][u][v] = s;
                  }
              for (int bb = 0; bb < nB; bb++)
                for (int a = 0; a < nA; a++) {
                  double s = 0.0;
                  for (int t = 0; t <= ca[a][0] + cb[bb][0]; t++)
                    for (int u = 0; u <= ca[a][1] + cb[bb][1]; u++)
                      for (int v = 0; v <= ca[a][2] + cb[bb][2]; v++)
                        s += bra.E[0][ca[a][0]][cb[bb][0]][t] * bra.E[1][ca[a][1]][cb[bb][1]][u] *
                             bra.E[2][ca[a][2]][cb[bb][2]][v] * G[t][u][v];
                  out[a + nA * (bb + nB * (c + nC * d))] += pref * s;
                }
            }
        }
    }
  for (int d = 0; d < nD; d++)
    for (int c = 0; c < nC; c++)
      for (int bb = 0; bb < nB; bb++)
        for (int a = 0; a < nA; a++)
          out[a + nA * (bb + nB * (c + nC * d))] *= nonaxial(la, ca[a]) * nonaxial(lb, cb[bb]) * nonaxial(lc, cc[c]) * nonaxial(ld, cd[d]);
}
void orc_md_eri_block(int nsh, const double *centers, const int *L, const int *nprim, const double *exps,
                      const double *coefs, int i, int j, int k, int l, double *out) {
  basis_t b; basis_init(&b, nsh, centers, L, nprim, exps, coefs);
  md_block(&b, i, j, k, l, out);
  basis_free(&b);
}

/* Schwarz bounds Q_ij = sqrt(max_{mu nu in ij} |(mu nu|mu nu)|), full nsh x nsh (symmetric) */
void orc_md_schwarz(int nsh, const double *centers, const int *L, const int *nprim, const double *exps,
                    const double *coefs, double *Q) {
  basis_t b; basis_init(&b, nsh, centers, L, nprim, exps, coefs);
#pragma omp parallel for schedule(dynamic)
  for (int i = 0; i < nsh; i++) {
    double *blk = (double *)malloc(sizeof(double) * 50625);
    for (int j = 0; j <= i; j++) {
      int nA = NCART(L[i]), nB = NCART(L[j]);
      md_block(&b, i, j, i, j, blk);
      double m = 0.0;
      for (int bb = 0; bb < nB; bb++)
        for (int a = 0; a < nA; a++) {
          double v = fabs(blk[a + nA * (bb + nB * (a + nA * bb))]);
          if (v > m) m = v;
        }
      Q[i + (long)nsh * j] = Q[j + (long)nsh * i] = sqrt(m);
    }
    free(blk);
  }
  basis_free(&b);
}

/* The integer screening rule shared (by specification, not by code) with the CUDA path:
 *   qlog(x)   = ceil( 8*log2(x) )  evaluated WITHOUT transcendental functions, so that CPU and GPU agree bit for
 *               bit:  x = m*2^e with m in [1,2)  ->  8e + #{ j in 0..7 : m > 2^(j/8) }   (-32768 for x < 1e-300)
 *   keep (ab|cd)  iff  qlog(Q_ab) + qlog(Q_cd) + dlog(ab,cd) >= tlog,
 *   dlog = max over the six shell blocks ab,cd,ac,ad,bc,bd of qlog(max|P_block|), tlog = qlog(tau).
 * (New functionality: the reference does not screen) */
int orc_qlog(double x) {
  static const double th[8] = {1.0, 1.0905077326652577, 1.189207115002721, 1.2968395546510096,
                               1.4142135623730951, 1.5422108254079407, 1.681792830507429, 1.8340080864093424};
  if (!(x >= 1e-300)) return -32768;
  uint64_t bits;
  memcpy(&bits, &x, 8);
  int e = (int)((bits >> 52) & 0x7ff) - 1023;
  uint64_t mb = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
  double m;
  memcpy(&m, &mb, 8);
  int c = 0;
  for (int j = 0; j < 8; j++) c += (m > th[j]) ? 1 : 0;
  return 8 * e + c;
}

/* fused J+K over unique quartets with screening; J,K exactly as the reference defines them
 * (J_mn = sum (mn|ls) P_ls ; K_mn = sum (ml|ns) P_ls) for symmetric P.  counts[0] = surviving shell
 * quartets, counts[1] = unique shell quartets visited.  Qin: optional nsh x nsh bounds (else computed).
 * rank/nranks: round-robin sharding of bra pairs for the multi-rank tests (partial J/K). */
void orc_md_jk(int nsh, const double *centers, const int *L, const int *nprim, const double *exps, const double *coefs,
               const double *P, double tau, const double *Qin, int rank, int nranks, double *J, double *K, long *counts) {
  basis_t b; basis_init(&b, nsh, centers, L, nprim, exps, coefs);
  long N = b.boff[nsh];
  double *Q = (double *)malloc(sizeof(double) * nsh * nsh);
  if (Qin) memcpy(Q, Qin, sizeof(double) * nsh * nsh);
  else orc_md_schwarz(nsh, centers, L, nprim, exps, coefs, Q);
  int *ql = (int *)malloc(sizeof(int) * nsh * nsh), *dl = (int *)malloc(sizeof(int) * nsh * nsh);
  for (int i = 0; i < nsh; i++)
    for (int j = 0; j < nsh; j++) {
      ql[i + nsh * j] = orc_qlog(Q[i + (long)nsh * j]);
      double m = 0.0;
      for (int a = b.boff[i]; a < b.boff[i + 1]; a++)
        for (int c = b.boff[j]; c < b.boff[j + 1]; c++) {
          double v = fabs(P[a + N * c]);
          if (v > m) m = v;
        }
      dl[i + nsh * j] = orc_qlog(m);
    }
  int tlog = orc_qlog(tau);
  memset(J, 0, sizeof(double) * N * N);
  memset(K, 0, sizeof(double) * N * N);
  long nsurv = 0, nvis = 0;
  long npair = (long)nsh * (nsh + 1) / 2;
#pragma omp parallel
  {
    double *Jl = (double *)calloc(N * N, sizeof(double)), *Kl = (double *)calloc(N * N, sizeof(double));
    double *blk = (double *)malloc(sizeof(double) * 50625);
    long ls = 0, lv = 0;
#pragma omp for schedule(dynamic, 4)
    for (long ij = 0; ij < npair; ij++) {
      if (ij % nranks != rank) continue;
      int i = (int)((sqrt(8.0 * ij + 1.0) - 1.0) / 2.0);
      while ((long)(i + 1) * (i + 2) / 2 <= ij) i++;
      while ((long)i * (i + 1) / 2 > ij) i--;
      int j = (int)(ij - (long)i * (i + 1) / 2);
      for (long kl = 0; kl <= ij; kl++) {
        int k = (int)((sqrt(8.0 * kl + 1.0) - 1.0) / 2.0);
        while ((long)(k + 1) * (k + 2) / 2 <= kl) k++;
        while ((long)k * (k + 1) / 2 > kl) k--;
        int l = (int)(kl - (long)k * (k + 1) / 2);
        lv++;
        int d = dl[i + nsh * j];
        if (dl[k + nsh * l] > d) d = dl[k + nsh * l];
        if (dl[i + nsh * k] > d) d = dl[i + nsh * k];
        if (dl[i + nsh * l] > d) d = dl[i + nsh * l];
        if (dl[j + nsh * k] > d) d = dl[j + nsh * k];
        if (dl[j + nsh * l] > d) d = dl[j + nsh * l];
        if ((long)ql[i + nsh * j] + ql[k + nsh * l] + d < tlog) continue;
        ls++;
        md_block(&b, i, j, k, l, blk);
        double deg = (i == j ? 1.0 : 2.0) * (k == l ? 1.0 : 2.0) * (ij == kl ? 1.0 : 2.0);
        int nA = NCART(L[i]), nB = NCART(L[j]), nC = NCART(L[k]), nD = NCART(L[l]);
        for (int dd = 0; dd < nD; dd++)
          for (int c = 0; c < nC; c++)
            for (int bb = 0; bb < nB; bb++)
              for (int a = 0; a < nA; a++) {
                double w = deg * blk[a + nA * (bb + nB * (c + nC * dd))];
                long fa = b.boff[i] + a, fb = b.boff[j] + bb, fc = b.boff[k] + c, fd = b.boff[l] + dd;
                Jl[fa + N * fb] += 0.5 * w * P[fc + N * fd];
                Jl[fc + N * fd] += 0.5 * w * P[fa + N * fb];
                Kl[fa + N * fc] += 0.25 * w * P[fb + N * fd];
                Kl[fa + N * fd] += 0.25 * w * P[fb + N * fc];
                Kl[fb + N * fc] += 0.25 * w * P[fa + N * fd];
                Kl[fb + N * fd] += 0.25 * w * P[fa + N * fc];
              }
      }
    }
#pragma omp critical
    {
      for (long t = 0; t < N * N; t++) { J[t] += Jl[t]; K[t] += Kl[t]; }
      nsurv += ls; nvis += lv;
    }
    free(Jl); free(Kl); free(blk);
  }
  /* symmetrise: X = (Xraw + Xraw^T)/2 */
  for (long a = 0; a < N; a++)
    for (long c = 0; c <= a; c++) {
      double v = 0.5 * (J[a + N * c] + J[c + N * a]); J[a + N * c] = J[c + N * a] = v;
      v = 0.5 * (K[a + N * c] + K[c + N * a]); K[a + N * c] = K[c + N * a] = v;
    }
  if (counts) { counts[0] = nsurv; counts[1] = nvis; }
  free(Q); free(ql); free(dl);
  basis_free(&b);
}

/* count only (no integrals): surviving / visited unique quartets for the integer rule above */
void orc_screen_count(int nsh, const int *L, const double *Q, const double *P, double tau, long *counts) {
  int *boff = (int *)malloc(sizeof(int) * (nsh + 1));
  boff[0] = 0;
  for (int i = 0; i < nsh; i++) boff[i + 1] = boff[i] + NCART(L[i]);
  long N = boff[nsh];
  int *ql = (int *)malloc(sizeof(int) * nsh * nsh), *dl = (int *)malloc(sizeof(int) * nsh * nsh);
  for (int i = 0; i < nsh; i++)
    for (int j = 0; j < nsh; j++) {
      ql[i + nsh * j] = orc_qlog(Q[i + (long)nsh * j]);
      double m = 0.0;
      for (int a = boff[i]; a < boff[i + 1]; a++)
        for (int c = boff[j]; c < boff[j + 1]; c++) {
          double v = fabs(P[a + N * c]);
          if (v > m) m = v;
        }
      dl[i + nsh * j] = orc_qlog(m);
    }
  int tlog = orc_qlog(tau);
  long nsurv = 0, nvis = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : nsurv, nvis)
  for (int i = 0; i < nsh; i++)
    for (int j = 0; j <= i; j++) {
      long ij = (long)i * (i + 1) / 2 + j;
      for (int k = 0; k <= i; k++)
        for (int l = 0; l <= k; l++) {
          long kl = (long)k * (k + 1) / 2 + l;
          if (kl > ij) continue;
          nvis++;
          int d = dl[i + nsh * j];
          if (dl[k + nsh * l] > d) d = dl[k + nsh * l];
          if (dl[i + nsh * k] > d) d = dl[i + nsh * k];
          if (dl[i + nsh * l] > d) d = dl[i + nsh * l];
          if (dl[j + nsh * k] > d) d = dl[j + nsh * k];
          if (dl[j + nsh * l] > d) d = dl[j + nsh * l];
          if ((long)ql[i + nsh * j] + ql[k + nsh * l] + d >= tlog) nsurv++;
        }
    }
  counts[0] = nsurv; counts[1] = nvis;
  free(ql); free(dl); free(boff);
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

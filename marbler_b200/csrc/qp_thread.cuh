// Barrier-certificate QP, one env per thread, everything in registers (small teams).
//
// Replaces rps create_single_integrator_barrier_certificate{,2} -> cvxopt.solvers.qp (called at
// /root/reference/robotarium_gym/utilities/controller.py:23).
// The iteration is cvxopt's coneqp for a pure 'l' cone (Mehrotra predictor-corrector, STEP 0.99,
// EXPON 3, default start, rps' options reltol = feastol = 1e-2, abstol 1e-7, maxiters 50) with
// cvxopt's data-dependent stopping rule, written in UNscaled variables: for the 'l' cone the
// Nesterov-Todd scaling W = diag(sqrt(s/z)) cancels algebraically, so (dx, ds, dz), the step
// lengths and sigma are the same numbers up to rounding.
//
// Structure that is exploited instead of a dense G (m x 2N):
//   row c = pair (i<j):  g_c = [ -a_c at block i, +a_c at block j ],  a_c = 2 (xi_i - xi_j) in R^2
//   (G v)_c = a_c . (v_j - v_i);  G'y scatters -+ a_c y_c;  K = 2I + G' diag(w) G is a weighted
//   graph Laplacian of 2x2 blocks  w_c a_c a_c'  -> packed lower-triangular Cholesky, no pivoting.
#pragma once
#include "common.cuh"

namespace mrb {

template <int N>
struct QpThread {
    static constexpr int n = 2 * N;
    static constexpr int m = N * (N - 1) / 2;
    static constexpr int KT = n * (n + 1) / 2;
    __device__ static constexpr int tri(int r, int c) { return r * (r + 1) / 2 + c; }   // r >= c

    double ax[m > 0 ? m : 1], ay[m > 0 ? m : 1], h[m > 0 ? m : 1];
    double L[KT], invd[n];

    // out[c] = (G v)_c
    __device__ __forceinline__ void G_mul(const double (&v)[n], double (&out)[m > 0 ? m : 1]) const
    {
        int c = 0;
#pragma unroll
        for (int i = 0; i < N - 1; i++)
#pragma unroll
            for (int j = i + 1; j < N; j++, c++)
                out[c] = ax[c] * (v[2 * j] - v[2 * i]) + ay[c] * (v[2 * j + 1] - v[2 * i + 1]);
    }
    // out += G' y
    __device__ __forceinline__ void GT_acc(const double (&y)[m > 0 ? m : 1], double (&out)[n]) const
    {
        int c = 0;
#pragma unroll
        for (int i = 0; i < N - 1; i++)
#pragma unroll
            for (int j = i + 1; j < N; j++, c++) {
                double tx = ax[c] * y[c], ty = ay[c] * y[c];
                out[2 * i] -= tx; out[2 * i + 1] -= ty;
                out[2 * j] += tx; out[2 * j + 1] += ty;
            }
    }
    // L := chol(2I + G' diag(w) G)   (lower, packed; invd = 1/diag)
    __device__ __forceinline__ void factor(const double (&w)[m > 0 ? m : 1])
    {
#pragma unroll
        for (int k = 0; k < KT; k++) L[k] = 0.0;
#pragma unroll
        for (int a = 0; a < n; a++) L[tri(a, a)] = 2.0;
        int c = 0;
#pragma unroll
        for (int i = 0; i < N - 1; i++)
#pragma unroll
            for (int j = i + 1; j < N; j++, c++) {
                double wx = w[c] * ax[c], wy = w[c] * ay[c];
                double pxx = wx * ax[c], pxy = wx * ay[c], pyy = wy * ay[c];
                L[tri(2 * i, 2 * i)] += pxx; L[tri(2 * i + 1, 2 * i)] += pxy; L[tri(2 * i + 1, 2 * i + 1)] += pyy;
                L[tri(2 * j, 2 * j)] += pxx; L[tri(2 * j + 1, 2 * j)] += pxy; L[tri(2 * j + 1, 2 * j + 1)] += pyy;
                L[tri(2 * j, 2 * i)] -= pxx; L[tri(2 * j, 2 * i + 1)] -= pxy;
                L[tri(2 * j + 1, 2 * i)] -= pxy; L[tri(2 * j + 1, 2 * i + 1)] -= pyy;
            }
        static_for<0, n>([&](auto J) {
            constexpr int j = decltype(J)::value;
            double d = L[tri(j, j)];
#pragma unroll
            for (int k = 0; k < j; k++) d -= L[tri(j, k)] * L[tri(j, k)];
            const double r = fast_rsqrt(d);
            invd[j] = r;
#pragma unroll
            for (int i = j + 1; i < n; i++) {
                double v = L[tri(i, j)];
#pragma unroll
                for (int k = 0; k < j; k++) v -= L[tri(i, k)] * L[tri(j, k)];
                L[tri(i, j)] = v * r;
            }
        });
    }
    __device__ __forceinline__ void solve(double (&b)[n]) const
    {
        static_for<0, n>([&](auto I) {
            constexpr int i = decltype(I)::value;
            double v = b[i];
#pragma unroll
            for (int k = 0; k < i; k++) v -= L[tri(i, k)] * b[k];
            b[i] = v * invd[i];
        });
        static_for<0, n>([&](auto I) {
            constexpr int i = n - 1 - decltype(I)::value;
            double v = b[i];
#pragma unroll
            for (int k = i + 1; k < n; k++) v -= L[tri(k, i)] * b[k];
            b[i] = v * invd[i];
        });
    }

    // xi: SI points; u: in = nominal dxi (already norm-limited to 0.15 by the position controller),
    // out = certified velocities.  Returns the number of interior-point iterations.
    __device__ __forceinline__ int run(const double (&xix)[N], const double (&xiy)[N], double (&ux)[N],
                                       double (&uy)[N], bool barrier_default)
    {
        double q[n], x[n];
#pragma unroll
        for (int i = 0; i < N; i++) {          // A.8: pre-clip columns of dxi to norm 0.2, f = -2 dxi
            double nrm = sqrt(ux[i] * ux[i] + uy[i] * uy[i]);
            if (nrm > kQpMagnitudeLimit) {
                double sc = kQpMagnitudeLimit / nrm;
                ux[i] *= sc; uy[i] *= sc;
            }
            q[2 * i] = -2.0 * ux[i];
            q[2 * i + 1] = -2.0 * uy[i];
        }
        if (m == 0) return 0;                  // single robot: u = dxi

        const double r2 = barrier_default ? 0.17 * 0.17 : 0.2 * 0.2;
        double hh = 0.0, qq = 0.0;
        {
            int c = 0;
#pragma unroll
            for (int i = 0; i < N - 1; i++)
#pragma unroll
                for (int j = i + 1; j < N; j++, c++) {
                    double ex = xix[i] - xix[j], ey = xiy[i] - xiy[j];
                    double hv = (ex * ex + ey * ey) - r2;
                    double gain = barrier_default ? 100.0 : (hv >= 0.0 ? 100.0 : 1e6);
                    h[c] = gain * (hv * hv * hv);
                    ax[c] = 2.0 * ex; ay[c] = 2.0 * ey;
                    hh += h[c] * h[c];
                }
#pragma unroll
            for (int a = 0; a < n; a++) qq += q[a] * q[a];
        }
        const double feas_x2 = 1e-4 * fmax(1.0, qq), feas_z2 = 1e-4 * fmax(1.0, hh);

        double s[m], z[m], t1[m], t2[m];
        // ---- default starting point: (2I + G'G) x = -q + G'h ; z = Gx - h ; s = -z ; shift
#pragma unroll
        for (int c = 0; c < m; c++) t1[c] = 1.0;
        factor(t1);
#pragma unroll
        for (int a = 0; a < n; a++) x[a] = -q[a];
        GT_acc(h, x);
        solve(x);
        G_mul(x, z);
        double ss = 0.0, ts = -INFINITY;
#pragma unroll
        for (int c = 0; c < m; c++) {
            z[c] -= h[c];
            s[c] = -z[c];
            ss += z[c] * z[c];
            ts = fmax(ts, z[c]);              // max(-s) = max(z)
        }
        const double nrm = fmax(sqrt(ss), 1.0);
        double tz = -ts;                       // placeholder, recomputed below
        tz = -INFINITY;
#pragma unroll
        for (int c = 0; c < m; c++) tz = fmax(tz, -z[c]);
        if (ts >= -1e-8 * nrm) {
#pragma unroll
            for (int c = 0; c < m; c++) s[c] += 1.0 + ts;
        }
        if (tz >= -1e-8 * nrm) {
#pragma unroll
            for (int c = 0; c < m; c++) z[c] += 1.0 + tz;
        }
        double gap = 0.0;
#pragma unroll
        for (int c = 0; c < m; c++) gap += s[c] * z[c];

        int iters = 0;
        for (; iters <= 50; iters++) {
            double rx[n], rz[m];
            // rx = 2x + q + G'z ; f0 = 1/2 x'Px + q'x ; rz = s + Gx - h
            double xq = 0.0, xrx = 0.0;
#pragma unroll
            for (int a = 0; a < n; a++) {
                rx[a] = 2.0 * x[a] + q[a];
                xrx += x[a] * rx[a];
                xq += x[a] * q[a];
            }
            const double f0 = 0.5 * (xrx + xq);
            GT_acc(z, rx);
            G_mul(x, rz);
            double resx = 0.0, resz = 0.0, zrz = 0.0;
#pragma unroll
            for (int a = 0; a < n; a++) resx += rx[a] * rx[a];
#pragma unroll
            for (int c = 0; c < m; c++) {
                rz[c] += s[c] - h[c];
                resz += rz[c] * rz[c];
                zrz += z[c] * rz[c];
            }
            const double pcost = f0, dcost = f0 + zrz - gap;
            bool gap_ok = gap <= 1e-7;
            if (pcost < 0.0) gap_ok = gap_ok || (gap <= -1e-2 * pcost);
            else if (dcost > 0.0) gap_ok = gap_ok || (gap <= 1e-2 * dcost);
            if ((resz <= feas_z2 && resx <= feas_x2 && gap_ok) || iters == 50) break;

            // w = z/s ; K = 2I + G' diag(w) G
            double w[m], sinv[m], zinv[m];
#pragma unroll
            for (int c = 0; c < m; c++) {
                sinv[c] = fast_rcp1(s[c]); zinv[c] = fast_rcp1(z[c]);
                w[c] = z[c] * sinv[c];
            }
            factor(w);

            // predictor: rc = -s.z  ->  K dx = -rx - G'((rc + z.rz)/s) = -rx - G'(w.rz - z)
            double dx[n], ds[m], dz[m];
#pragma unroll
            for (int c = 0; c < m; c++) t1[c] = z[c] - w[c] * rz[c];
#pragma unroll
            for (int a = 0; a < n; a++) dx[a] = -rx[a];
            GT_acc(t1, dx);
            solve(dx);
            G_mul(dx, ds);
            double dsdz = 0.0, tmax = 0.0;
#pragma unroll
            for (int c = 0; c < m; c++) {
                ds[c] = -rz[c] - ds[c];
                dz[c] = -z[c] - w[c] * ds[c];
                t2[c] = ds[c] * dz[c];                     // Mehrotra correction term
                dsdz += t2[c];
                tmax = fmax(tmax, fmax(-ds[c] * sinv[c], -dz[c] * zinv[c]));
            }
            double step = tmax <= 1.0 ? 1.0 : fast_rcp(tmax);
            double sg = fmin(1.0, fmax(0.0, 1.0 - step + dsdz * fast_rcp(gap) * (step * step)));
            const double sigmamu = sg * sg * sg * (gap / m);

            // corrector: rc = -s.z - ds_aff.dz_aff + sigma mu
#pragma unroll
            for (int c = 0; c < m; c++) {
                t2[c] = (sigmamu - t2[c]) * sinv[c];       // (rc + s.z)/s
                t1[c] = z[c] - w[c] * rz[c] - t2[c];       // -(rc + z.rz)/s
            }
#pragma unroll
            for (int a = 0; a < n; a++) dx[a] = -rx[a];
            GT_acc(t1, dx);
            solve(dx);
            G_mul(dx, ds);
            tmax = 0.0;
#pragma unroll
            for (int c = 0; c < m; c++) {
                ds[c] = -rz[c] - ds[c];
                dz[c] = t2[c] - z[c] - w[c] * ds[c];       // (rc - z.ds)/s
                tmax = fmax(tmax, fmax(-ds[c] * sinv[c], -dz[c] * zinv[c]));
            }
            step = tmax <= 0.99 ? 1.0 : 0.99 * fast_rcp(tmax);
#pragma unroll
            for (int a = 0; a < n; a++) x[a] += step * dx[a];
            gap = 0.0;
#pragma unroll
            for (int c = 0; c < m; c++) {
                s[c] += step * ds[c];
                z[c] += step * dz[c];
                gap += s[c] * z[c];
            }
        }
#pragma unroll
        for (int i = 0; i < N; i++) {
            ux[i] = x[2 * i];
            uy[i] = x[2 * i + 1];
        }
        return iters;
    }
};

}  // namespace mrb

// Barrier-certificate QP, one env per thread, DUAL (constraint-space) normal equations.
//
// Same iteration as qp_thread.cuh (cvxopt coneqp for a pure 'l' cone with rps' options, SURVEY.md
// App. A.8 / A.9; called by the reference at utilities/controller.py:23) but the Newton system is
// reduced to the m = N(N-1)/2 multipliers instead of the 2N velocities: with P = 2I
//      (1/2 G G' + diag(s/z)) dz = rz - 1/2 G rx + rc/z ,   dx = -1/2 (rx + G' dz) ,   ds = (rc - s.dz)/z
// which is algebraically the same step.  For teams of up to 4 robots m <= 6 < 2N = 8, so the Cholesky
// factor is 6x6 (21 numbers) instead of 8x8 (36): fewer flops, shorter dependency chains and fewer live
// registers.  G G' is never stored: entry (c,d) is +-(a_c . a_d) when the two pairs share a robot.
#pragma once
#include "common.cuh"

namespace mrb {

template <int N>
struct QpDual {
    static constexpr int n = 2 * N;
    static constexpr int m = N * (N - 1) / 2;
    static constexpr int MT = m * (m + 1) / 2;
    __device__ static constexpr int tri(int r, int c) { return r * (r + 1) / 2 + c; }   // r >= c

    double ax[m], ay[m], h[m];
    double L[MT], invd[m];

    __device__ __forceinline__ void G_mul(const double (&v)[n], double (&out)[m]) const
    {
        int c = 0;
#pragma unroll
        for (int i = 0; i < N - 1; i++)
#pragma unroll
            for (int j = i + 1; j < N; j++, c++)
                out[c] = ax[c] * (v[2 * j] - v[2 * i]) + ay[c] * (v[2 * j + 1] - v[2 * i + 1]);
    }
    __device__ __forceinline__ void GT_acc(const double (&y)[m], double (&out)[n]) const
    {
        int c = 0;
#pragma unroll
        for (int i = 0; i < N - 1; i++)
#pragma unroll
            for (int j = i + 1; j < N; j++, c++) {
                const double tx = ax[c] * y[c], ty = ay[c] * y[c];
                out[2 * i] -= tx; out[2 * i + 1] -= ty;
                out[2 * j] += tx; out[2 * j + 1] += ty;
            }
    }
    // L := chol(1/2 G G' + diag(d))
    __device__ __forceinline__ void factor(const double (&d)[m])
    {
        {
            int c = 0;
#pragma unroll
            for (int i = 0; i < N - 1; i++)
#pragma unroll
                for (int j = i + 1; j < N; j++, c++) {
                    L[tri(c, c)] = fma(ax[c], ax[c], ay[c] * ay[c]) + d[c];      // 1/2 |g_c|^2 = |a_c|^2
                    int e = 0;
#pragma unroll
                    for (int k = 0; k < N - 1; k++)
#pragma unroll
                        for (int l = k + 1; l < N; l++, e++) {
                            if (e < c) {
                                const int sgn = (i == k) + (j == l) - (i == l) - (j == k);
                                if (sgn == 0) L[tri(c, e)] = 0.0;
                                else {
                                    const double dot = fma(ax[c], ax[e], ay[c] * ay[e]);
                                    L[tri(c, e)] = sgn > 0 ? 0.5 * dot : -0.5 * dot;
                                }
                            }
                        }
                }
        }
        static_for<0, m>([&](auto J) {
            constexpr int j = decltype(J)::value;
            double dj = L[tri(j, j)];
#pragma unroll
            for (int k = 0; k < j; k++) dj = fma(-L[tri(j, k)], L[tri(j, k)], dj);
            const double r = fast_rsqrt(dj);
            invd[j] = r;
#pragma unroll
            for (int i = j + 1; i < m; i++) {
                double v = L[tri(i, j)];
#pragma unroll
                for (int k = 0; k < j; k++) v = fma(-L[tri(i, k)], L[tri(j, k)], v);
                L[tri(i, j)] = v * r;
            }
        });
    }
    // b := (L L')^-1 b
    __device__ __forceinline__ void solve(double (&b)[m]) const
    {
        static_for<0, m>([&](auto I) {
            constexpr int i = decltype(I)::value;
            double v = b[i];
#pragma unroll
            for (int k = 0; k < i; k++) v = fma(-L[tri(i, k)], b[k], v);
            b[i] = v * invd[i];
        });
        static_for<0, m>([&](auto I) {
            constexpr int i = m - 1 - decltype(I)::value;
            double v = b[i];
#pragma unroll
            for (int k = i + 1; k < m; k++) v = fma(-L[tri(k, i)], b[k], v);
            b[i] = v * invd[i];
        });
    }

    __device__ __forceinline__ int run(const double (&xix)[N], const double (&xiy)[N], double (&ux)[N],
                                       double (&uy)[N], bool barrier_default)
    {
        double q[n], x[n];
#pragma unroll
        for (int i = 0; i < N; i++) {          // A.8: pre-clip columns of dxi to norm 0.2, f = -2 dxi
            const double n2 = ux[i] * ux[i] + uy[i] * uy[i];
            if (n2 > kQpMagnitudeLimit * kQpMagnitudeLimit) {
                const double sc = kQpMagnitudeLimit / sqrt(n2);
                ux[i] *= sc; uy[i] *= sc;
            }
            q[2 * i] = -2.0 * ux[i];
            q[2 * i + 1] = -2.0 * uy[i];
        }
        const double r2 = barrier_default ? 0.17 * 0.17 : 0.2 * 0.2;
        double hh = 0.0, qq = 0.0;
        {
            int c = 0;
#pragma unroll
            for (int i = 0; i < N - 1; i++)
#pragma unroll
                for (int j = i + 1; j < N; j++, c++) {
                    const double ex = xix[i] - xix[j], ey = xiy[i] - xiy[j];
                    const double hv = (ex * ex + ey * ey) - r2;
                    const double gain = barrier_default ? 100.0 : (hv >= 0.0 ? 100.0 : 1e6);
                    h[c] = gain * (hv * hv * hv);
                    ax[c] = 2.0 * ex; ay[c] = 2.0 * ey;
                    hh = fma(h[c], h[c], hh);
                }
        }
#pragma unroll
        for (int a = 0; a < n; a++) qq = fma(q[a], q[a], qq);
        // (feastol * max(1, |q|))^2 and (feastol * max(1, |h|))^2
        const double feas_x2 = 1e-4 * fmax(1.0, qq), feas_z2 = 1e-4 * fmax(1.0, hh);

        double s[m], z[m], t1[m], t2[m];
        // ---- default starting point [P G'; G -I][x; z] = [-q; h]:  (1/2 GG' + I) z = -h - 1/2 G q,  x = -1/2 (q + G'z)
#pragma unroll
        for (int c = 0; c < m; c++) t1[c] = 1.0;
        factor(t1);
        G_mul(q, z);
#pragma unroll
        for (int c = 0; c < m; c++) z[c] = -h[c] - 0.5 * z[c];
        solve(z);
#pragma unroll
        for (int a = 0; a < n; a++) x[a] = q[a];
        GT_acc(z, x);
#pragma unroll
        for (int a = 0; a < n; a++) x[a] *= -0.5;
        G_mul(x, z);                           // z = Gx - h as cvxopt forms it
        double ss = 0.0, ts = -INFINITY, tz = -INFINITY;
#pragma unroll
        for (int c = 0; c < m; c++) {
            z[c] -= h[c];
            s[c] = -z[c];
            ss = fma(z[c], z[c], ss);
            ts = fmax(ts, z[c]);
            tz = fmax(tz, -z[c]);
        }
        const double nrm = fmax(sqrt(ss), 1.0);
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
        for (int c = 0; c < m; c++) gap = fma(s[c], z[c], gap);

        int iters = 0;
        for (; iters <= 50; iters++) {
            double rx[n], rz[m];
            double xq = 0.0, xrx = 0.0;
#pragma unroll
            for (int a = 0; a < n; a++) {
                rx[a] = fma(2.0, x[a], q[a]);
                xrx = fma(x[a], rx[a], xrx);
                xq = fma(x[a], q[a], xq);
            }
            const double f0 = 0.5 * (xrx + xq);
            GT_acc(z, rx);
            G_mul(x, rz);
            double resx = 0.0, resz = 0.0, zrz = 0.0;
#pragma unroll
            for (int a = 0; a < n; a++) resx = fma(rx[a], rx[a], resx);
#pragma unroll
            for (int c = 0; c < m; c++) {
                rz[c] += s[c] - h[c];
                resz = fma(rz[c], rz[c], resz);
                zrz = fma(z[c], rz[c], zrz);
            }
            const double pcost = f0, dcost = f0 + zrz - gap;
            // cvxopt's stopping rule without sqrt / division: relgap <= reltol  <=>  gap <= 1e-2 * denominator,
            // pres = sqrt(resz)/resz0 <= feastol  <=>  resz <= (1e-2 * resz0)^2   (equal up to 1 ulp at the threshold)
            bool gap_ok = gap <= 1e-7;
            if (pcost < 0.0) gap_ok = gap_ok || (gap <= -1e-2 * pcost);
            else if (dcost > 0.0) gap_ok = gap_ok || (gap <= 1e-2 * dcost);
            if ((resz <= feas_z2 && resx <= feas_x2 && gap_ok) || iters == 50) break;

            double zinv[m], sinv[m], dz[m], ds[m];
#pragma unroll
            for (int c = 0; c < m; c++) {
                zinv[c] = fast_rcp1(z[c]);
                sinv[c] = fast_rcp1(s[c]);
                t1[c] = s[c] * zinv[c];                     // D = s/z
            }
            factor(t1);
            // base right-hand side shared by predictor and corrector: rz - 1/2 G rx - s
            G_mul(rx, t2);
#pragma unroll
            for (int c = 0; c < m; c++) t2[c] = rz[c] - 0.5 * t2[c] - s[c];
            // predictor (rc = -s.z)
#pragma unroll
            for (int c = 0; c < m; c++) dz[c] = t2[c];
            solve(dz);
            double dsdz = 0.0, tmax = 0.0;
#pragma unroll
            for (int c = 0; c < m; c++) {
                ds[c] = -s[c] - t1[c] * dz[c];
                const double p = ds[c] * dz[c];
                dsdz += p;
                tmax = fmax(tmax, fmax(-ds[c] * sinv[c], -dz[c] * zinv[c]));
                ds[c] = p;                                    // keep only the Mehrotra correction term
            }
            double step = tmax <= 1.0 ? 1.0 : fast_rcp(tmax);          // t == 0 ? 1 : min(1, 1/t)
            const double sg = fmin(1.0, fmax(0.0, 1.0 - step + dsdz * fast_rcp(gap) * (step * step)));
            const double sigmamu = sg * sg * sg * (gap / m);
            // corrector (rc = -s.z - ds_aff.dz_aff + sigma mu)
            double corr[m];
#pragma unroll
            for (int c = 0; c < m; c++) {
                corr[c] = (sigmamu - ds[c]) * zinv[c];        // (rc + s.z)/z
                dz[c] = t2[c] + corr[c];
            }
            solve(dz);
            tmax = 0.0;
#pragma unroll
            for (int c = 0; c < m; c++) {
                ds[c] = corr[c] - s[c] - t1[c] * dz[c];
                tmax = fmax(tmax, fmax(-ds[c] * sinv[c], -dz[c] * zinv[c]));
            }
            step = tmax <= 0.99 ? 1.0 : 0.99 * fast_rcp(tmax);          // t == 0 ? 1 : min(1, 0.99/t)
            // dx = -1/2 (rx + G'dz)
            GT_acc(dz, rx);
            const double hstep = -0.5 * step;
#pragma unroll
            for (int a = 0; a < n; a++) x[a] = fma(hstep, rx[a], x[a]);
            gap = 0.0;
#pragma unroll
            for (int c = 0; c < m; c++) {
                s[c] = fma(step, ds[c], s[c]);
                z[c] = fma(step, dz[c], z[c]);
                gap = fma(s[c], z[c], gap);
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

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
#include "qp_store.cuh"

namespace mrb {

// per-env vectors of the dual solver that may live in the shared store (bit i of VMASK <-> DualVec i); lengths m or n
enum DualVec { DV_H = 0, DV_RZ, DV_T2, DV_Q, DV_RX, DV_CORR, DV_P, DV_D, DV_S, DV_Z, DV_AX, DV_AY, DV_X, DV_COUNT };

// LS: 1 = the strictly-lower part of the m x m Cholesky factor lives in the shared store as well (the reciprocals
// of its diagonal stay in registers).  TPB: threads per CTA (the interleave stride of the store).
template <int N, int TPB = 1, unsigned VMASK = 0, int LS = 0>
struct QpDual {
    static constexpr int n = 2 * N;
    static constexpr int m = N * (N - 1) / 2;
    static constexpr int MT = m * (m - 1) / 2;                 // strictly-lower entries
    __device__ static constexpr int low(int r, int c) { return r * (r - 1) / 2 + c; }   // r > c

    __device__ static constexpr bool in_smem(int id) { return (VMASK >> id) & 1u; }
    __device__ static constexpr int len(int id) { return (id == DV_Q || id == DV_RX || id == DV_X) ? n : m; }
    __device__ static constexpr int offset(int id) { int k = 0; for (int b = 0; b < id; b++) k += in_smem(b) ? len(b) : 0; return k; }
    static constexpr int kVecDoubles = offset(DV_COUNT);
    static constexpr int kStoreDoubles = kVecDoubles + (LS ? MT : 0);      // doubles per thread
    template <int ID> using Vec = MVec<len(ID), TPB, in_smem(ID)>;

    Vec<DV_AX> ax;
    Vec<DV_AY> ay;
    MVec<MT, TPB, LS != 0> L;
    double invd[m];
    double *Vs;

    __device__ __forceinline__ QpDual(double *Vs_ = nullptr)
        : ax(Vs_, offset(DV_AX)), ay(Vs_, offset(DV_AY)), L(Vs_, kVecDoubles), Vs(Vs_) {}

    // sink(c, (G v)_c)
    template <typename V, typename F>
    __device__ __forceinline__ void G_mul(V &&v, F &&sink) const
    {
        int c = 0;
#pragma unroll
        for (int i = 0; i < N - 1; i++)
#pragma unroll
            for (int j = i + 1; j < N; j++, c++)
                sink(c, ax.get(c) * (v(2 * j) - v(2 * i)) + ay.get(c) * (v(2 * j + 1) - v(2 * i + 1)));
    }
    // out += G' y
    template <typename F>
    __device__ __forceinline__ void GT_acc(F &&y, double (&out)[n]) const
    {
        int c = 0;
#pragma unroll
        for (int i = 0; i < N - 1; i++)
#pragma unroll
            for (int j = i + 1; j < N; j++, c++) {
                const double yc = y(c);
                const double tx = ax.get(c) * yc, ty = ay.get(c) * yc;
                out[2 * i] -= tx; out[2 * i + 1] -= ty;
                out[2 * j] += tx; out[2 * j + 1] += ty;
            }
    }
    // entry (c, e), e < c, of 1/2 G G': +-(a_c . a_e)/2 when the two pairs share a robot, else 0
    template <int C, int E>
    __device__ __forceinline__ double gram() const
    {
        constexpr int ci = pair_i(C), cj = pair_j(C), ei = pair_i(E), ej = pair_j(E);
        constexpr int sgn = (ci == ei) + (cj == ej) - (ci == ej) - (cj == ei);
        if constexpr (sgn == 0) return 0.0;
        else {
            const double dot = fma(ax.get(C), ax.get(E), ay.get(C) * ay.get(E));
            return sgn > 0 ? 0.5 * dot : -0.5 * dot;
        }
    }
    __device__ static constexpr int pair_i(int c) { int i = 0; while (c >= N - 1 - i) { c -= N - 1 - i; i++; } return i; }
    __device__ static constexpr int pair_j(int c) { int i = 0; while (c >= N - 1 - i) { c -= N - 1 - i; i++; } return i + 1 + c; }

    // L := chol(1/2 G G' + diag(d)), column by column (left-looking): pivot k from its own row, then the entries below
    // it, which are independent of each other.  The Gram entries are formed where they are consumed (never stored).
    // Per column the dependent chain is (last product of the sums) -> rsqrt -> scale, 72 cycles, against a chain that
    // grows with the row index in a row-by-row order (measured latencies: profiles/r02_fp64_latency_microbench.txt).
    template <typename D>
    __device__ __forceinline__ void factor(D &&d)
    {
        static_for<0, m>([&](auto K_) {
            constexpr int k = decltype(K_)::value;
            const double a = ax.get(k), b = ay.get(k);
            double dk = fma(a, a, b * b) + d(k);                                    // 1/2 |g_k|^2 = |a_k|^2
#pragma unroll
            for (int j = 0; j < k; j++) { const double l = L.get(low(k, j)); dk = fma(-l, l, dk); }
            const double r = fast_rsqrt(dk);
            invd[k] = r;
            static_for<k + 1, m>([&](auto I_) {
                constexpr int i = decltype(I_)::value;
                double v = gram<i, k>();
#pragma unroll
                for (int j = 0; j < k; j++) v = fma(-L.get(low(i, j)), L.get(low(k, j)), v);
                L.set(low(i, k), v * r);
            });
        });
    }
    // b := (L L')^-1 b
    __device__ __forceinline__ void solve(double (&b)[m]) const
    {
        static_for<0, m>([&](auto I) {
            constexpr int i = decltype(I)::value;
            double v = b[i];
#pragma unroll
            for (int k = 0; k < i; k++) v = fma(-L.get(low(i, k)), b[k], v);
            b[i] = v * invd[i];
        });
        // backward sweep in row order: once x_i is known every earlier row receives -L_ik x_i
        static_for<0, m>([&](auto I) {
            constexpr int i = m - 1 - decltype(I)::value;
            const double xi = b[i] * invd[i];
            b[i] = xi;
#pragma unroll
            for (int k = 0; k < i; k++) b[k] = fma(-L.get(low(i, k)), xi, b[k]);
        });
    }

    __device__ __forceinline__ int run(const double (&xix)[N], const double (&xiy)[N], double (&ux)[N],
                                       double (&uy)[N], bool barrier_default)
    {
        Vec<DV_Q> q(Vs, offset(DV_Q));
        Vec<DV_X> x(Vs, offset(DV_X));
#pragma unroll
        for (int i = 0; i < N; i++) {          // A.8: pre-clip columns of dxi to norm 0.2, f = -2 dxi
            const double n2 = ux[i] * ux[i] + uy[i] * uy[i];
            if (n2 > kQpMagnitudeLimit * kQpMagnitudeLimit) {
                const double sc = kQpMagnitudeLimit / sqrt(n2);
                ux[i] *= sc; uy[i] *= sc;
            }
            q.set(2 * i, -2.0 * ux[i]);
            q.set(2 * i + 1, -2.0 * uy[i]);
        }
        Vec<DV_H> h(Vs, offset(DV_H));
        Vec<DV_S> s(Vs, offset(DV_S));
        Vec<DV_Z> z(Vs, offset(DV_Z));
        Vec<DV_RZ> rz(Vs, offset(DV_RZ));
        Vec<DV_RX> rxs(Vs, offset(DV_RX));
        Vec<DV_T2> t2(Vs, offset(DV_T2));
        Vec<DV_D> dd(Vs, offset(DV_D));
        Vec<DV_CORR> corr(Vs, offset(DV_CORR));
        Vec<DV_P> pp(Vs, offset(DV_P));

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
                    const double hc = gain * (hv * hv * hv);
                    h.set(c, hc);
                    ax.set(c, 2.0 * ex); ay.set(c, 2.0 * ey);
                    hh = fma(hc, hc, hh);
                }
        }
        {
            double q0 = 0.0, q1 = 0.0;
#pragma unroll
            for (int a = 0; a < n; a += 2) {
                const double qa = q.get(a), qb = q.get(a + 1);
                q0 = fma(qa, qa, q0); q1 = fma(qb, qb, q1);
            }
            qq = q0 + q1;
        }
        // (feastol * max(1, |q|))^2 and (feastol * max(1, |h|))^2
        const double feas_x2 = 1e-4 * dmax(qq, 1.0), feas_z2 = 1e-4 * dmax(hh, 1.0);

        // ---- default starting point [P G'; G -I][x; z] = [-q; h]:  (1/2 GG' + I) z = -h - 1/2 G q,  x = -1/2 (q + G'z)
        factor([](int) { return 1.0; });
        double tv[m];
        G_mul([&](int a) { return q.get(a); }, [&](int c, double g) { tv[c] = -h.get(c) - 0.5 * g; });
        solve(tv);
        {
            double xa[n];
#pragma unroll
            for (int a = 0; a < n; a++) xa[a] = q.get(a);
            GT_acc([&](int c) { return tv[c]; }, xa);
#pragma unroll
            for (int a = 0; a < n; a++) x.set(a, xa[a] * -0.5);
        }
        double ss = 0.0, ts = -INFINITY, tz = -INFINITY;
        G_mul([&](int a) { return x.get(a); }, [&](int c, double g) {       // z = Gx - h as cvxopt forms it
            const double zc = g - h.get(c);
            z.set(c, zc);
            ss = fma(zc, zc, ss);
            ts = dmax(zc, ts);
            tz = dmax(-zc, tz);
        });
        const double nrm = ss > 1.0 ? ss * fast_rsqrt(ss) : 1.0;          // max(|z|, 1)
        const bool do_s = ts >= -1e-8 * nrm, do_z = tz >= -1e-8 * nrm;
        double gap = 0.0;
#pragma unroll
        for (int c = 0; c < m; c++) {
            const double z0 = z.get(c);
            double sc = -z0, zc = z0;
            if (do_s) sc += 1.0 + ts;
            if (do_z) zc += 1.0 + tz;
            s.set(c, sc); z.set(c, zc);
            gap = fma(sc, zc, gap);
        }

        int iters = 0;
        for (; iters <= 50; iters++) {
            double rx[n];
            // the sums of this block are accumulated in two halves each (even / odd terms): half the dependent depth
            double xq = 0.0, xrx = 0.0, xq1 = 0.0, xrx1 = 0.0;
#pragma unroll
            for (int a = 0; a < n; a += 2) {
                const double xa = x.get(a), qa = q.get(a), xb = x.get(a + 1), qb = q.get(a + 1);
                rx[a] = fma(2.0, xa, qa);
                rx[a + 1] = fma(2.0, xb, qb);
                xrx = fma(xa, rx[a], xrx); xrx1 = fma(xb, rx[a + 1], xrx1);
                xq = fma(xa, qa, xq); xq1 = fma(xb, qb, xq1);
            }
            const double f0 = 0.5 * ((xrx + xrx1) + (xq + xq1));
            GT_acc([&](int c) { return z.get(c); }, rx);
            double resx = 0.0, resx1 = 0.0, resz = 0.0, resz1 = 0.0, zrz = 0.0, zrz1 = 0.0;
#pragma unroll
            for (int a = 0; a < n; a += 2) { resx = fma(rx[a], rx[a], resx); resx1 = fma(rx[a + 1], rx[a + 1], resx1); }
            resx += resx1;
            G_mul([&](int a) { return x.get(a); }, [&](int c, double g) {
                const double r = g + (s.get(c) - h.get(c));
                rz.set(c, r);
                if (c & 1) { resz1 = fma(r, r, resz1); zrz1 = fma(z.get(c), r, zrz1); }
                else { resz = fma(r, r, resz); zrz = fma(z.get(c), r, zrz); }
            });
            resz += resz1; zrz += zrz1;
            const double pcost = f0, dcost = f0 + zrz - gap;
            // cvxopt's stopping rule without sqrt / division: relgap <= reltol  <=>  gap <= 1e-2 * denominator,
            // pres = sqrt(resz)/resz0 <= feastol  <=>  resz <= (1e-2 * resz0)^2   (equal up to 1 ulp at the threshold)
            bool gap_ok = gap <= 1e-7;
            if (pcost < 0.0) gap_ok = gap_ok || (gap <= -1e-2 * pcost);
            else if (dcost > 0.0) gap_ok = gap_ok || (gap <= 1e-2 * dcost);
            if ((resz <= feas_z2 && resx <= feas_x2 && gap_ok) || iters == 50) break;

            // 1/s and 1/z are recomputed where they are used (same function of the same input: identical bits)
            auto inv_s = [&](int c) { return fast_rcp1(s.get(c)); };
            auto inv_z = [&](int c) { return fast_rcp1(z.get(c)); };
#pragma unroll
            for (int c = 0; c < m; c++) dd.set(c, s.get(c) * inv_z(c));           // D = s/z
            factor([&](int c) { return dd.get(c); });
            // base right-hand side shared by predictor and corrector: rz - 1/2 G rx - s
            G_mul([&](int a) { return rx[a]; }, [&](int c, double g) { t2.set(c, rz.get(c) - 0.5 * g - s.get(c)); });
#pragma unroll
            for (int a = 0; a < n; a++) rxs.set(a, rx[a]);                          // rx is needed again after the solves
            // predictor (rc = -s.z)
            double dz[m];
#pragma unroll
            for (int c = 0; c < m; c++) dz[c] = t2.get(c);
            solve(dz);
            double dsdz = 0.0, dsdz1 = 0.0, tmax;
            double cand[m];
#pragma unroll
            for (int c = 0; c < m; c++) {
                const double sc = s.get(c);
                const double dsc = -sc - dd.get(c) * dz[c];
                const double p = dsc * dz[c];
                if (c & 1) dsdz1 += p; else dsdz += p;
                cand[c] = dmax(-dsc * inv_s(c), -dz[c] * inv_z(c));
                pp.set(c, p);                                  // keep only the Mehrotra correction term
            }
            dsdz += dsdz1;
            tmax = dmax(tree_max(cand), 0.0);
            double step = tmax <= 1.0 ? 1.0 : fast_rcp(tmax);          // t == 0 ? 1 : min(1, 1/t)
            const double sg = dmin(dmax(1.0 - step + dsdz * fast_rcp(gap) * (step * step), 0.0), 1.0);
            const double sigmamu = sg * sg * sg * (gap / m);
            // corrector (rc = -s.z - ds_aff.dz_aff + sigma mu)
#pragma unroll
            for (int c = 0; c < m; c++) {
                const double cr = (sigmamu - pp.get(c)) * inv_z(c);        // (rc + s.z)/z
                corr.set(c, cr);
                dz[c] = t2.get(c) + cr;
            }
            solve(dz);
#pragma unroll
            for (int c = 0; c < m; c++) {
                const double dsc = corr.get(c) - s.get(c) - dd.get(c) * dz[c];
                cand[c] = dmax(-dsc * inv_s(c), -dz[c] * inv_z(c));
            }
            tmax = dmax(tree_max(cand), 0.0);
            step = tmax <= 0.99 ? 1.0 : 0.99 * fast_rcp(tmax);          // t == 0 ? 1 : min(1, 0.99/t)
            // dx = -1/2 (rx + G'dz)
            double rx2[n];
#pragma unroll
            for (int a = 0; a < n; a++) rx2[a] = rxs.get(a);
            GT_acc([&](int c) { return dz[c]; }, rx2);
            const double hstep = -0.5 * step;
#pragma unroll
            for (int a = 0; a < n; a++) x.set(a, fma(hstep, rx2[a], x.get(a)));
            gap = 0.0;
            double gap1 = 0.0;
#pragma unroll
            for (int c = 0; c < m; c++) {
                const double s0 = s.get(c);
                const double dsc = corr.get(c) - s0 - dd.get(c) * dz[c];
                const double sc = fma(step, dsc, s0), zc = fma(step, dz[c], z.get(c));
                s.set(c, sc); z.set(c, zc);
                if (c & 1) gap1 = fma(sc, zc, gap1); else gap = fma(sc, zc, gap);
            }
            gap += gap1;
        }
#pragma unroll
        for (int i = 0; i < N; i++) {
            ux[i] = x.get(2 * i);
            uy[i] = x.get(2 * i + 1);
        }
        return iters;
    }
};

}  // namespace mrb

/* ORACLE / TEST INFRASTRUCTURE ONLY -- see marbler_oracle.h.  PARITY UNPINNED vs real rps/cvxopt.
 *
 * Scalar float64 restatement of the MARBLER env step.  Every function cites the reference
 * (/root/reference/robotarium_gym/...) or, for the un-vendored packages, SURVEY.md Appendix A and
 * the restated module under oracle/shims it follows.  Deliberately literal: dense G, dense
 * Cholesky, Nesterov-Todd scaled iterates, sin/cos/atan2 every sub-step -- the CUDA product is
 * free to be cleverer, this file is not.
 */
#include "marbler_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

#define NMAX ORC_NMAX
#define PMAX ORC_PMAX
#define QN (2 * NMAX)                    /* QP variables */
#define QM (NMAX * (NMAX - 1) / 2)       /* QP rows */

/* rps RobotariumABC constants (oracle/shims/rps/robotarium_abc.py) */
static const double TIME_STEP = 0.033;
static const double MAX_LIN = 0.2;
static const double BX0 = -1.6, BY0 = -1.0, BW = 3.2, BH = 2.0;
/* controller constants (App. A.6-A.8) */
static const double PROJ = 0.05, SI_VEL_LIMIT = 0.15, ANG_LIMIT = M_PI, QP_MAG_LIMIT = 0.2;

int orc_max_threads(void) { return 1; }   /* scalar port; callers thread over env slices (ctypes drops the GIL) */

/* ------------------------------------------------------------------ layout */
static int scen_nf(const orc_config *c)
{
    if (c->scenario == ORC_PCP) return 2 * c->num_prey;
    if (c->scenario == ORC_SIMPLE) return 2;
    return 0;
}
static int scen_ni(const orc_config *c)
{
    switch (c->scenario) {
    case ORC_PCP: return 2 * c->num_prey;
    case ORC_WAREHOUSE: return c->N;
    case ORC_MATERIAL: return c->N + 6;
    case ORC_ARCTIC: return 96 + 1 + 2 * c->N;
    default: return 0;
    }
}
int orc_nf(const orc_config *c) { return 6 * c->N + scen_nf(c); }
int orc_ni(const orc_config *c) { return 3 + scen_ni(c); }
int orc_obs_dim(const orc_config *c)
{
    int N = c->N, K = c->num_neighbors >= N - 1 ? N - 1 : c->num_neighbors;
    switch (c->scenario) {
    case ORC_PCP: return (c->capability_aware ? 6 : 4) * (c->num_neighbors + 1);   /* PredatorCapturePrey.py:52 */
    case ORC_WAREHOUSE: (void)K; return 3 * (c->num_neighbors + 1);                 /* warehouse.py:70 */
    case ORC_MATERIAL: return c->capability_aware ? 11 : 9;                         /* MaterialTransport.py:55-58 */
    case ORC_ARCTIC: return 30;                                                     /* ArcticTransport.py:19 */
    default: return 2 * (N + 1);                                                    /* simple.py:98 */
    }
}

/* ------------------------------------------------------------------ dense Cholesky helpers */
static int chol_factor(int n, double *A, int lda)       /* lower, in place; 0 ok, 1 not PD */
{
    for (int j = 0; j < n; j++) {
        double d = A[j * lda + j];
        for (int k = 0; k < j; k++) d -= A[j * lda + k] * A[j * lda + k];
        if (!(d > 0.0)) return 1;
        d = sqrt(d);
        A[j * lda + j] = d;
        for (int i = j + 1; i < n; i++) {
            double v = A[i * lda + j];
            for (int k = 0; k < j; k++) v -= A[i * lda + k] * A[j * lda + k];
            A[i * lda + j] = v / d;
        }
    }
    return 0;
}
static void chol_solve(int n, const double *L, int lda, double *b)
{
    for (int i = 0; i < n; i++) {
        double v = b[i];
        for (int k = 0; k < i; k++) v -= L[i * lda + k] * b[k];
        b[i] = v / L[i * lda + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double v = b[i];
        for (int k = i + 1; k < n; k++) v -= L[k * lda + i] * b[k];
        b[i] = v / L[i * lda + i];
    }
}
static double dotn(int n, const double *a, const double *b)
{
    double s = 0.0;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* ------------------------------------------------------------------ cvxopt coneqp, 'l' cone only
 * Follows oracle/shims/cvxopt/solvers.py:coneqp_l statement by statement:
 * default start, NT scaling d = sqrt(s/z), lmbda = sqrt(s*z), kkt 'chol2', Mehrotra with
 * STEP 0.99 / EXPON 3, rps options reltol = feastol = 1e-2, abstol 1e-7, maxiters 50.
 * P = 2 I.  G is dense m x n row-major.  Returns iterations; x holds the returned iterate. */
typedef struct {
    double G[QM * QN], Gs[QM * QN], S[QN * QN];
    double h[QM], s[QM], z[QM], d[QM], lm[QM], lmsq[QM], rz[QM], bs[QM], bz[QM], uz[QM], us[QM],
        dsdz_aff[QM], ds[QM], dz[QM];
    double q[QN], x[QN], rx[QN], bx[QN], dx[QN];
} qp_ws;

static int coneqp_l(qp_ws *w, int n, int m)
{
    const int MAXITERS = 50;
    const double ABSTOL = 1e-7, RELTOL = 1e-2, FEASTOL = 1e-2, STEP = 0.99;
    double *G = w->G, *h = w->h, *q = w->q, *x = w->x, *s = w->s, *z = w->z;

    double resx0 = fmax(1.0, sqrt(dotn(n, q, q)));
    double resz0 = fmax(1.0, sqrt(dotn(m, h, h)));

    /* start: (P + G'G) x = -q + G'h ; z = Gx - h ; s = -z ; shift into the cone */
    for (int a = 0; a < n; a++) {
        for (int b = 0; b <= a; b++) {
            double v = (a == b) ? 2.0 : 0.0;
            for (int r = 0; r < m; r++) v += G[r * n + a] * G[r * n + b];
            w->S[a * n + b] = v;
        }
        double v = -q[a];
        for (int r = 0; r < m; r++) v += G[r * n + a] * h[r];
        x[a] = v;
    }
    chol_factor(n, w->S, n);
    chol_solve(n, w->S, n, x);
    for (int r = 0; r < m; r++) {
        z[r] = dotn(n, G + r * n, x) - h[r];
        s[r] = -z[r];
    }
    double nrms = sqrt(dotn(m, s, s)), ts = -INFINITY;
    for (int r = 0; r < m; r++) ts = fmax(ts, -s[r]);
    if (ts >= -1e-8 * fmax(nrms, 1.0))
        for (int r = 0; r < m; r++) s[r] += 1.0 + ts;
    double nrmz = sqrt(dotn(m, z, z)), tz = -INFINITY;
    for (int r = 0; r < m; r++) tz = fmax(tz, -z[r]);
    if (tz >= -1e-8 * fmax(nrmz, 1.0))
        for (int r = 0; r < m; r++) z[r] += 1.0 + tz;

    double gap = dotn(m, s, z);

    for (int iters = 0; iters <= MAXITERS; iters++) {
        /* residuals, costs, stopping rule */
        double xq = dotn(n, x, q), xrx = 0.0;
        for (int a = 0; a < n; a++) {
            w->rx[a] = 2.0 * x[a] + q[a];
            xrx += x[a] * w->rx[a];
        }
        double f0 = 0.5 * (xrx + xq);
        for (int a = 0; a < n; a++) {
            double v = 0.0;
            for (int r = 0; r < m; r++) v += G[r * n + a] * z[r];
            w->rx[a] += v;
        }
        double resx = sqrt(dotn(n, w->rx, w->rx));
        for (int r = 0; r < m; r++) w->rz[r] = s[r] + dotn(n, G + r * n, x) - h[r];
        double resz = sqrt(dotn(m, w->rz, w->rz));
        double pcost = f0, dcost = f0 + dotn(m, z, w->rz) - gap;
        int have_relgap = 1;
        double relgap = 0.0;
        if (pcost < 0.0) relgap = gap / -pcost;
        else if (dcost > 0.0) relgap = gap / dcost;
        else have_relgap = 0;
        double pres = resz / resz0, dres = resx / resx0;
        if ((pres <= FEASTOL && dres <= FEASTOL && (gap <= ABSTOL || (have_relgap && relgap <= RELTOL)))
            || iters == MAXITERS)
            return iters;

        if (iters == 0)
            for (int r = 0; r < m; r++) {
                w->d[r] = sqrt(s[r] / z[r]);
                w->lm[r] = sqrt(s[r] * z[r]);
            }
        for (int r = 0; r < m; r++) w->lmsq[r] = w->lm[r] * w->lm[r];

        /* chol2: S = P + Gs'Gs, Gs = diag(1/d) G */
        for (int r = 0; r < m; r++)
            for (int a = 0; a < n; a++) w->Gs[r * n + a] = G[r * n + a] / w->d[r];
        for (int a = 0; a < n; a++)
            for (int b = 0; b <= a; b++) {
                double v = (a == b) ? 2.0 : 0.0;
                for (int r = 0; r < m; r++) v += w->Gs[r * n + a] * w->Gs[r * n + b];
                w->S[a * n + b] = v;
            }
        if (chol_factor(n, w->S, n)) return iters;     /* cvxopt: status 'unknown', last iterate */

        double mu = gap / m, sigma = 0.0, dsdz = 0.0, step = 1.0;
        for (int i = 0; i < 2; i++) {
            for (int r = 0; r < m; r++) {
                double bs = -w->lmsq[r] + sigma * mu;
                if (i == 1) bs -= w->dsdz_aff[r];
                bs /= w->lm[r];                                   /* f4_no_ir: sinv */
                w->bs[r] = bs;
                w->bz[r] = (-w->rz[r] - w->d[r] * bs) / w->d[r];
            }
            for (int a = 0; a < n; a++) {
                double v = -w->rx[a];
                for (int r = 0; r < m; r++) v += w->Gs[r * n + a] * w->bz[r];
                w->dx[a] = v;
            }
            chol_solve(n, w->S, n, w->dx);
            for (int r = 0; r < m; r++) {
                w->uz[r] = dotn(n, w->Gs + r * n, w->dx) - w->bz[r];   /* scaled dz */
                w->us[r] = w->bs[r] - w->uz[r];                          /* scaled ds */
            }
            if (i == 0) {
                dsdz = dotn(m, w->us, w->uz);
                for (int r = 0; r < m; r++) w->dsdz_aff[r] = w->us[r] * w->uz[r];
            }
            double tsm = -INFINITY, tzm = -INFINITY;
            for (int r = 0; r < m; r++) {
                w->ds[r] = w->us[r] / w->lm[r];
                w->dz[r] = w->uz[r] / w->lm[r];
                tsm = fmax(tsm, -w->ds[r]);
                tzm = fmax(tzm, -w->dz[r]);
            }
            double t = fmax(0.0, fmax(tsm, tzm));
            if (t == 0.0) step = 1.0;
            else if (i == 0) step = fmin(1.0, 1.0 / t);
            else step = fmin(1.0, STEP / t);
            if (i == 0) {
                double sg = fmin(1.0, fmax(0.0, 1.0 - step + dsdz / gap * (step * step)));
                sigma = sg * sg * sg;
            }
        }
        for (int a = 0; a < n; a++) x[a] += step * w->dx[a];
        gap = 0.0;
        for (int r = 0; r < m; r++) {                              /* misc.update_scaling, 'l' block */
            double a = sqrt((1.0 + step * w->ds[r]) * w->lm[r]);
            double b = sqrt((1.0 + step * w->dz[r]) * w->lm[r]);
            w->d[r] = w->d[r] * a / b;
            w->lm[r] = a * b;
            s[r] = w->d[r] * w->lm[r];
            z[r] = w->lm[r] / w->d[r];
            gap += w->lm[r] * w->lm[r];
        }
    }
    return MAXITERS;
}

/* rps create_single_integrator_barrier_certificate{,2} (App. A.8; shims/rps/utilities/
 * barrier_certificates.py:_solve); MARBLER selects them at utilities/controller.py:13-16 */
int orc_barrier_qp(int N, int barrier_default, const double *dxi_in, const double *xi, double *u)
{
    static _Thread_local qp_ws *tls_ws = NULL;          /* one workspace per host thread, never freed */
    if (!tls_ws) tls_ws = (qp_ws *)malloc(sizeof(qp_ws));
    qp_ws *w = tls_ws;
    int n = 2 * N, m = N * (N - 1) / 2, count = 0;
    double r = barrier_default ? 0.17 : 0.2;
    double dxi[2 * NMAX];
    memcpy(dxi, dxi_in, sizeof(double) * 2 * N);
    memset(w->G, 0, sizeof(double) * (size_t)(m > 0 ? m : 1) * n);
    for (int i = 0; i < N - 1; i++)
        for (int j = i + 1; j < N; j++) {
            double ex = xi[i] - xi[j], ey = xi[N + i] - xi[N + j];
            double hh = (ex * ex + ey * ey) - pow(r, 2);
            w->G[count * n + 2 * i] = -2 * ex;
            w->G[count * n + 2 * i + 1] = -2 * ey;
            w->G[count * n + 2 * j] = 2 * ex;
            w->G[count * n + 2 * j + 1] = 2 * ey;
            double gain = barrier_default ? 100.0 : (hh >= 0 ? 100.0 : 1e6);
            w->h[count] = gain * pow(hh, 3);
            count++;
        }
    for (int i = 0; i < N; i++) {
        double nrm = sqrt(dxi[i] * dxi[i] + dxi[N + i] * dxi[N + i]);
        if (nrm > QP_MAG_LIMIT) {
            dxi[i] *= QP_MAG_LIMIT / nrm;
            dxi[N + i] *= QP_MAG_LIMIT / nrm;
        }
        w->q[2 * i] = -2 * dxi[i];
        w->q[2 * i + 1] = -2 * dxi[N + i];
    }
    int iters = 0;
    if (m == 0) {
        for (int a = 0; a < n; a++) w->x[a] = -w->q[a] / 2.0;
    } else {
        iters = coneqp_l(w, n, m);
    }
    for (int i = 0; i < N; i++) {
        u[i] = w->x[2 * i];
        u[N + i] = w->x[2 * i + 1];
    }
    return iters;
}

/* Controller.set_velocities (utilities/controller.py:20-25) + Robotarium.set_velocities saturation
 * (App. A.2), as called from roboEnv.py:64-65 */
int orc_controller(int N, int barrier_default, const double *pose, const double *goal, double *dxu)
{
    double xi[2 * NMAX], dxi[2 * NMAX], u[2 * NMAX];
    for (int i = 0; i < N; i++) {                       /* uni_to_si_states (A.7) */
        xi[i] = pose[i] + PROJ * cos(pose[2 * N + i]);
        xi[N + i] = pose[N + i] + PROJ * sin(pose[2 * N + i]);
    }
    for (int i = 0; i < N; i++) {                       /* si_position_controller (A.6) */
        dxi[i] = 1 * (goal[i] - xi[i]);
        dxi[N + i] = 1 * (goal[N + i] - xi[N + i]);
        double nrm = sqrt(dxi[i] * dxi[i] + dxi[N + i] * dxi[N + i]);
        if (nrm > SI_VEL_LIMIT) {
            dxi[i] *= SI_VEL_LIMIT / nrm;
            dxi[N + i] *= SI_VEL_LIMIT / nrm;
        }
    }
    int iters = orc_barrier_qp(N, barrier_default, dxi, xi, u);
    const double max_ang = 2 * (0.016 / 0.11) * (MAX_LIN / 0.016);
    for (int i = 0; i < N; i++) {                       /* si_to_uni_dyn (A.7) */
        double cs = cos(pose[2 * N + i]), ss = sin(pose[2 * N + i]);
        double v = cs * u[i] + ss * u[N + i];
        double om = (1 / PROJ) * (-ss * u[i] + cs * u[N + i]);
        if (om > ANG_LIMIT) om = ANG_LIMIT;
        if (om < -ANG_LIMIT) om = -ANG_LIMIT;
        if (fabs(v) > MAX_LIN) v = MAX_LIN * (v > 0 ? 1.0 : -1.0);      /* A.2 */
        if (fabs(om) > max_ang) om = max_ang * (om > 0 ? 1.0 : -1.0);
        dxu[i] = v;
        dxu[N + i] = om;
    }
    return iters;
}

/* ------------------------------------------------------------------ goal generation
 * every scenario's Agent.generate_goal (PredatorCapturePrey/agent.py:48-76, warehouse.py:19-45,
 * MaterialTransport.py:19-46, ArcticTransport/agent.py:89-137, simple.py:32-60) */
static void generate_goal(const orc_config *c, double step, int action, double *gx, double *gy)
{
    double x = *gx, y = *gy;
    double cx = x < c->LEFT ? c->LEFT : (x > c->RIGHT ? c->RIGHT : x);
    double cy = y < c->UP ? c->UP : (y > c->DOWN ? c->DOWN : y);
    switch (action) {
    case 0: *gx = fmax(x - step, c->LEFT); *gy = cy; break;
    case 1: *gx = fmin(x + step, c->RIGHT); *gy = cy; break;
    case 2: *gx = cx; *gy = fmax(y - step, c->UP); break;
    case 3: *gx = cx; *gy = fmin(y + step, c->DOWN); break;
    default: *gx = cx; *gy = cy; break;
    }
}

static double agent_step(const orc_config *c, int i, const int32_t *pixel_type)
{
    if (c->scenario == ORC_MATERIAL)                                  /* MaterialTransport.py:71-74 */
        return i < c->n_fast ? c->fast_step : c->slow_step;
    if (c->scenario == ORC_ARCTIC) {                                  /* ArcticTransport/agent.py:94-112 */
        if (i < 2) return c->fast_step;                               /* drones */
        int p = pixel_type[i];
        if (i == 3) return p == 1 ? c->slow_step : (p == 2 ? c->fast_step : c->step_dist);   /* water */
        return p == 1 ? c->fast_step : (p == 2 ? c->slow_step : c->step_dist);              /* ice */
    }
    return c->step_dist;
}

static double norm2(double a, double b) { return sqrt(a * a + b * b); }

/* utilities/misc.py:20-25 get_nearest_neighbors; canonical order = ascending distance, ties by
 * lower index (SURVEY section 7 hard part 7) */
static void nearest_neighbors(int N, const double *pose, int agent, int K, int *out)
{
    double dist[NMAX];
    int used[NMAX];
    for (int x = 0; x < N; x++) {
        dist[x] = norm2(pose[x] - pose[agent], pose[N + x] - pose[N + agent]);
        used[x] = (x == agent);
    }
    for (int k = 0; k < K; k++) {
        int best = -1;
        for (int x = 0; x < N; x++)
            if (!used[x] && (best < 0 || dist[x] < dist[best])) best = x;
        used[best] = 1;
        out[k] = best;
    }
}
static int neighbor_list(const orc_config *c, const double *pose, int agent, int *out)
{
    int N = c->N;
    if (c->num_neighbors >= N - 1) {                     /* PredatorCapturePrey.py:198-199 */
        int k = 0;
        for (int x = 0; x < N; x++)
            if (x != agent) out[k++] = x;
        return N - 1;
    }
    nearest_neighbors(N, pose, agent, c->num_neighbors, out);
    return c->num_neighbors;
}

/* ArcticTransport.py:136-143 */
static void at_cell(double x, double y, int *row, int *col)
{
    int r = -(int)((y - 1) / .25), cc = (int)((x + 1.5) / .25);
    *row = r < 0 ? 0 : (r > 7 ? 7 : r);
    *col = cc < 0 ? 0 : (cc > 11 ? 11 : cc);
}

/* ------------------------------------------------------------------ the env step
 * <Scenario>.step -> roboEnv.step (utilities/roboEnv.py:38-96) -> scenario tail; SURVEY App. C */
void orc_step(const orc_config *c, double *sf, int32_t *si, const int32_t *actions,
              double *obs, double *reward, double *dist, int32_t *out_i)
{
    const int N = c->N, D = orc_obs_dim(c);
    double *pose = sf, *prev = sf + 3 * N, *scf = sf + 6 * N;
    int32_t *sci = si + 3;
    double goal[2 * NMAX], vel[2 * NMAX];
    int msg = 0, n_qp = 0, n_it = 0, max_it = 0;

    si[0] += 1;                                          /* episode_steps, first line of every step() */

    /* roboEnv.py:42 -> _generate_step_goal_positions: goal from the pose at entry */
    const int32_t *pixel_type = sci + 97;
    for (int i = 0; i < N; i++) {
        int a = actions[i];
        if (c->scenario == ORC_MATERIAL) a = a / 4;      /* MaterialTransport.py:23 */
        goal[i] = pose[i];
        goal[N + i] = pose[N + i];
        generate_goal(c, agent_step(c, i, pixel_type), a, &goal[i], &goal[N + i]);
        dist[i] = 0.0;
        vel[i] = vel[N + i] = 0.0;
    }

    for (int k = 0; k < c->update_frequency; k++) {      /* roboEnv.py:52 */
        if (si[1])                                       /* :55-56 */
            for (int i = 0; i < N; i++) dist[i] += norm2(pose[i] - prev[i], pose[N + i] - prev[N + i]);
        memcpy(prev, pose, sizeof(double) * 3 * N);      /* :59 */
        si[1] = 1;
        if (k % c->ctrl_period == 0 || c->robotarium) {  /* :63-65 */
            int it = orc_controller(N, c->barrier_default, pose, goal, vel);
            n_it += it;
            if (it > max_it) max_it = it;
            n_qp++;
        }
        /* Robotarium.step (A.3): validate on the entering pose (A.4), then integrate in place */
        int viol_b = 0, viol_c = 0;
        for (int i = 0; i < N; i++) {
            double x = pose[i], y = pose[N + i];
            if (x < BX0 || x > (BX0 + BW) || y < BY0 || y > (BY0 + BH)) viol_b = 1;
        }
        for (int j = 0; j < N - 1; j++)
            for (int l = j + 1; l < N; l++) {
                double off = c->collision_offset;           /* first_position / second_position of _validate */
                double x1 = pose[j] + off * cos(pose[2 * N + j]), y1 = pose[N + j] + off * sin(pose[2 * N + j]);
                double x2 = pose[l] + off * cos(pose[2 * N + l]), y2 = pose[N + l] + off * sin(pose[2 * N + l]);
                if (norm2(x1 - x2, y1 - y2) <= c->collision_diameter) viol_c = 1;
            }
        for (int i = 0; i < N; i++) {
            double th = pose[2 * N + i];
            pose[i] = pose[i] + TIME_STEP * cos(th) * vel[i];
            pose[N + i] = pose[N + i] + TIME_STEP * sin(th) * vel[i];
            th = th + TIME_STEP * vel[N + i];
            pose[2 * N + i] = atan2(sin(th), cos(th));
        }
        if (c->penalize_violations && (viol_c || viol_b)) {     /* roboEnv.py:82-94 */
            msg = viol_c * 1 + viol_b * 2;
            for (int i = 0; i < N; i++) dist[i] += norm2(pose[i] - prev[i], pose[N + i] - prev[N + i]);
            break;
        }
    }

    int done = 0, remaining = 0;
    memset(obs, 0, sizeof(double) * N * D);

    if (c->scenario == ORC_PCP) {
        const int P = c->num_prey, od = c->capability_aware ? 6 : 4;
        double *prey = scf;
        int32_t *sensed = sci, *captured = sci + P;
        int unseen0 = P, left0 = P;
        for (int p = 0; p < P; p++) { unseen0 -= sensed[p]; left0 -= captured[p]; }
        /* _update_tracking_and_locations (PredatorCapturePrey.py:72-95) */
        for (int p = 0; p < P; p++) {
            if (captured[p]) continue;
            if (!sensed[p])
                for (int a = 0; a < N; a++) {
                    double rad = a < c->num_predators ? c->predator_radius : 0.0;
                    if (norm2(pose[a] - prey[2 * p], pose[N + a] - prey[2 * p + 1]) <= rad) { sensed[p] = 1; break; }
                }
            if (sensed[p])
                for (int a = 0; a < N; a++) {
                    double rad = a < c->num_predators ? 0.0 : c->capture_radius;
                    if (actions[a] == 4 && norm2(pose[a] - prey[2 * p], pose[N + a] - prey[2 * p + 1]) <= rad) {
                        captured[p] = 1;
                        break;
                    }
                }
        }
        int unseen = P, left = P;
        for (int p = 0; p < P; p++) { unseen -= sensed[p]; left -= captured[p]; }
        /* Agent.get_observation (agent.py:19-46) */
        double blk[NMAX][6];
        for (int a = 0; a < N; a++) {
            double srad = a < c->num_predators ? c->predator_radius : 0.0;
            double crad = a < c->num_predators ? 0.0 : c->capture_radius;
            double closest = -1, px = -5, py = -5;
            for (int p = 0; p < P; p++) {
                if (captured[p]) continue;
                double dd = norm2(pose[a] - prey[2 * p], pose[N + a] - prey[2 * p + 1]);
                if (dd <= srad && (dd < closest || closest == -1)) { px = prey[2 * p]; py = prey[2 * p + 1]; closest = dd; }
            }
            blk[a][0] = pose[a]; blk[a][1] = pose[N + a]; blk[a][2] = px; blk[a][3] = py;
            blk[a][4] = srad; blk[a][5] = crad;
        }
        for (int a = 0; a < N; a++) {                     /* get_observations (:178-207) */
            int nb[NMAX], K = neighbor_list(c, pose, a, nb);
            memcpy(obs + a * D, blk[a], sizeof(double) * od);
            for (int k = 0; k < K; k++) memcpy(obs + a * D + (k + 1) * od, blk[nb[k]], sizeof(double) * od);
        }
        double r;
        if (msg) { r = c->violation_reward; done = 1; }    /* :155-159 */
        else {
            r = 0;
            r += (unseen0 - unseen) * c->sense_reward;     /* get_rewards (:209-216) */
            r += (left0 - left) * c->capture_reward;
            r += c->time_penalty;
            done = (si[0] > c->max_episode_steps) || left == 0;
        }
        for (int a = 0; a < N; a++) reward[a] = r;
        remaining = left;
    } else if (c->scenario == ORC_WAREHOUSE) {
        int32_t *loaded = sci;
        double blk[NMAX][3];
        for (int a = 0; a < N; a++) { blk[a][0] = pose[a]; blk[a][1] = pose[N + a]; blk[a][2] = loaded[a]; }
        for (int a = 0; a < N; a++) {                     /* get_observations (warehouse.py:124-143) */
            int nb[NMAX], K = neighbor_list(c, pose, a, nb);
            memcpy(obs + a * D, blk[a], sizeof(double) * 3);
            for (int k = 0; k < K; k++) memcpy(obs + a * D + (k + 1) * 3, blk[nb[k]], sizeof(double) * 3);
        }
        if (msg) {
            for (int a = 0; a < N; a++) reward[a] = c->violation_reward;
            done = 1;
        } else {
            for (int a = 0; a < N; a++) {                 /* get_rewards (:145-178); even index = Green (:63-65) */
                double x = pose[a], y = pose[N + a], r = 0;
                int green = (a % 2 == 0);
                if (loaded[a]) {
                    if (x < -1.5 + c->goal_width && ((green && y > 0) || (!green && y <= 0))) { r = c->unload_reward; loaded[a] = 0; }
                } else {
                    if (x > 1.5 - c->goal_width && ((!green && y > 0) || (green && y <= 0))) { r = c->load_reward; loaded[a] = 1; }
                }
                reward[a] = r;
            }
            done = si[0] > c->max_episode_steps;
        }
    } else if (c->scenario == ORC_MATERIAL) {
        int32_t *load = sci, *zone = sci + N, *messages = sci + N + 2;
        for (int i = 0; i < 4; i++) messages[i] = actions[i] % 4;          /* MaterialTransport.py:119-120 */
        for (int a = 0; a < N; a++) {                     /* get_observations (:150-159) */
            double *o = obs + a * D;
            o[0] = pose[a]; o[1] = pose[N + a]; o[2] = load[a]; o[3] = zone[0]; o[4] = zone[1];
            for (int i = 0; i < 4; i++) o[5 + i] = messages[i];
            if (c->capability_aware) {
                o[9] = a < c->n_fast ? c->small_torque : c->large_torque;
                o[10] = a < c->n_fast ? c->fast_step : c->slow_step;
            }
        }
        double r;
        if (msg) { r = c->violation_reward; done = 1; }
        else {
            r = c->time_penalty;                          /* get_reward (:161-189) */
            for (int a = 0; a < N; a++) {
                int torque = a < c->n_fast ? c->small_torque : c->large_torque;
                double x = pose[a], y = pose[N + a];
                if (load[a] > 0) {
                    if (x < -1.5 + c->goal_width) { r += load[a] * c->unload_reward; load[a] = 0; }
                } else {
                    int zi = -1;
                    if (x > 1.5 - c->goal_width) zi = 1;
                    else if (norm2(x - 0, y - 0) <= c->zone1_radius) zi = 0;
                    if (zi >= 0) {
                        if (zone[zi] > torque) { load[a] = torque; zone[zi] -= torque; }
                        else { load[a] = zone[zi]; zone[zi] = 0; }
                        r += load[a] * c->load_reward;
                    }
                }
            }
            done = si[0] > c->max_episode_steps;
            if (!done) {
                done = zone[0] == 0 && zone[1] == 0;
                for (int a = 0; a < N && done; a++)
                    if (load[a] != 0) done = 0;
            }
        }
        for (int a = 0; a < N; a++) reward[a] = r;
        remaining = zone[0] + zone[1];
        for (int a = 0; a < N; a++) remaining += load[a];
    } else if (c->scenario == ORC_ARCTIC) {
        int32_t *grid = sci, goal_col = sci[96], *ptype = sci + 97, *reached = sci + 97 + N;
        int row[NMAX], col[NMAX];
        for (int i = 0; i < N; i++) at_cell(pose[i], pose[N + i], &row[i], &col[i]);
        double gx = goal_col * .25 - 1.5, gy = (-1 * .25 + .75);              /* get_pose_from_cell([1, g]) */
        static const int perm[4][3] = {{1, 2, 3}, {0, 2, 3}, {3, 0, 1}, {2, 0, 1}};   /* agent.py:42-69 */
        for (int a = 0; a < N; a++) {                     /* Agent.get_observation (agent.py:14-87) */
            double *o = obs + a * D;
            int k = 0;
            ptype[a] = grid[row[a] * 12 + col[a]];
            if (ptype[a] == 3) reached[a] = 1;
            o[k++] = pose[a]; o[k++] = pose[N + a]; o[k++] = ptype[a];
            for (int t = 0; t < 3; t++) {
                int b = perm[a][t];
                o[k++] = pose[b]; o[k++] = pose[N + b]; o[k++] = grid[row[b] * 12 + col[b]];
            }
            o[k++] = gx; o[k++] = gy;
            for (int i = 0; i < 2; i++) {
                int left = col[i] > 0 ? col[i] - 1 : col[i], right = col[i] < 11 ? col[i] + 1 : col[i];
                int up = row[i] > 0 ? row[i] - 1 : row[i], down = row[i] < 7 ? row[i] + 1 : row[i];
                o[k++] = grid[up * 12 + left]; o[k++] = grid[row[i] * 12 + left]; o[k++] = grid[down * 12 + left];
                o[k++] = grid[up * 12 + col[i]]; o[k++] = grid[down * 12 + col[i]];
                o[k++] = grid[up * 12 + right]; o[k++] = grid[row[i] * 12 + right]; o[k++] = grid[down * 12 + right];
            }
        }
        double r;
        if (msg) { r = c->violation_reward; done = 1; }
        else {
            r = 0;                                         /* get_reward (ArcticTransport.py:125-134) */
            for (int a = 2; a < N; a++) {
                if (!reached[a]) r += c->not_reached_penalty;
                if (ptype[a] != 3) {
                    double dd = norm2(pose[a] - gx, pose[N + a] - gy);
                    r += c->dist_multiplier * (dd * dd);
                }
            }
            done = si[0] > c->max_episode_steps;
            if (!done) {
                done = 1;
                for (int a = 2; a < N; a++)
                    if (!reached[a]) done = 0;
            }
        }
        for (int a = 0; a < N; a++) reward[a] = r;
    } else {                                               /* Simple (simple.py:155-225) */
        double *g = scf;
        for (int a = 0; a < N; a++) {
            double *o = obs + a * D;
            int k = 0;
            o[k++] = pose[a]; o[k++] = pose[N + a];
            for (int b = 0; b < N; b++)
                if (b != a) { o[k++] = pose[b]; o[k++] = pose[N + b]; }
            o[k++] = g[0]; o[k++] = g[1];
        }
        if (msg) {
            for (int a = 0; a < N; a++) reward[a] = c->violation_reward;
            done = 1;
        } else {
            for (int a = 0; a < N; a++) {
                double ddx = pose[a] - g[0], ddy = pose[N + a] - g[1];
                double r = -(ddx * ddx + ddy * ddy);
                reward[a] = r * c->reward_scaler;
            }
            done = si[0] > c->max_episode_steps;
        }
    }
    out_i[0] = msg; out_i[1] = done; out_i[2] = remaining; out_i[3] = n_qp; out_i[4] = n_it; out_i[5] = max_it;
}

/* ------------------------------------------------------------------ reset (distributional parity)
 * Philox4x32-10 (Salmon et al., SC'11) keyed by the seed, counter = (env id, episode, block).
 * The draw procedure is this repo's own (the reference uses numpy's legacy global RandomState and
 * Python's `random`); only the DISTRIBUTIONS follow the reference. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct { uint32_t ctr[4], key[2], buf[4]; int have; } rng_t;
static void rng_init(rng_t *g, uint64_t seed, uint64_t env_id, uint32_t episode)
{
    g->key[0] = (uint32_t)seed; g->key[1] = (uint32_t)(seed >> 32);
    g->ctr[0] = (uint32_t)env_id; g->ctr[1] = (uint32_t)(env_id >> 32);
    g->ctr[2] = episode; g->ctr[3] = 0; g->have = 0;
}
static uint32_t rng_u32(rng_t *g)
{
    if (!g->have) { orc_philox4x32_10(g->ctr, g->key, g->buf); g->ctr[3]++; g->have = 4; }
    return g->buf[4 - g->have--];
}
static uint32_t rng_below(rng_t *g, uint32_t n) { return (uint32_t)(((uint64_t)rng_u32(g) * n) >> 32); }
static double rng_unit(rng_t *g)                      /* 53-bit uniform in [0,1) */
{
    uint32_t a = rng_u32(g) >> 5, b = rng_u32(g) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}
static double rng_normal(rng_t *g)                    /* Box-Muller, one variate per two uniforms */
{
    double u1 = 1.0 - rng_unit(g), u2 = rng_unit(g);
    return sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
}

/* rps generate_initial_conditions (App. A.5) + utilities/misc.py:49-63: N distinct cells of an
 * xr x yr grid, uniformly, in order; cell -> (ix, iy) = divmod(cell, yr) */
static void spawn_grid(const orc_spawn *sp, rng_t *g, double *x, double *y, double *th, int stride)
{
    uint64_t taken = 0;
    int cells = sp->xr * sp->yr;
    for (int i = 0; i < sp->count; i++) {
        int r = (int)rng_below(g, (uint32_t)(cells - i)), cell = 0;
        for (cell = 0; cell < cells; cell++) {
            if (taken >> cell & 1) continue;
            if (r-- == 0) break;
        }
        taken |= 1ull << cell;
        int ix = cell / sp->yr, iy = cell % sp->yr;
        x[i * stride] = ((ix * sp->spacing - sp->w2) + sp->sx1) + sp->sx2;
        y[i * stride] = ((iy * sp->spacing - sp->h2) + sp->sy1) + sp->sy2;
        if (th) {
            double t = 0.0;
            if (sp->random_theta) {                        /* warehouse.py:93 keeps rps' random heading */
                t = rng_unit(g) * 2 * M_PI - M_PI;
                t = atan2(sin(t), cos(t));                 /* the zero-velocity step of roboEnv.py:112 */
            }
            th[i * stride] = t;
        }
    }
}

void orc_reset(const orc_config *c, uint64_t seed, uint64_t env_id, double *sf, int32_t *si)
{
    const int N = c->N;
    double *pose = sf, *prev = sf + 3 * N, *scf = sf + 6 * N;
    int32_t *sci = si + 3;
    rng_t g;
    rng_init(&g, seed, env_id, (uint32_t)si[2]);
    si[0] = 0; si[1] = 0; si[2] += 1;
    memset(prev, 0, sizeof(double) * 3 * N);
    memset(sci, 0, sizeof(int32_t) * scen_ni(c));
    if (c->scenario == ORC_ARCTIC) {                       /* ArcticTransport.py:28-33, 56-82 */
        static const double sx[4] = {-.3, .3, -.9, .9};
        for (int i = 0; i < N; i++) { pose[i] = sx[i]; pose[N + i] = -.8; pose[2 * N + i] = atan2(sin(M_PI / 2), cos(M_PI / 2)); }
        for (int k = 0; k < 96; k++) sci[k] = (int32_t)rng_below(&g, 3);
        int gc = 1 + (int)rng_below(&g, 11);
        sci[gc] = sci[gc - 1] = sci[12 + gc] = sci[12 + gc - 1] = 3;
        for (int k = 1; k < 11; k++) sci[7 * 12 + k] = 0;
        sci[96] = gc;
        return;
    }
    spawn_grid(&c->spawn_robots, &g, pose, pose + N, pose + 2 * N, 1);
    if (c->scenario == ORC_PCP) spawn_grid(&c->spawn_other, &g, scf, scf + 1, NULL, 2);     /* PredatorCapturePrey.py:128-130 */
    if (c->scenario == ORC_SIMPLE) spawn_grid(&c->spawn_other, &g, scf, scf + 1, NULL, 2);  /* simple.py:141-144 */
    if (c->scenario == ORC_MATERIAL)                        /* MaterialTransport.py:99-100 */
        for (int k = 0; k < 2; k++) sci[N + k] = (int32_t)(c->zone_mu[k] + c->zone_sigma[k] * rng_normal(&g));
}

void orc_reset_batch(const orc_config *c, int64_t B, uint64_t seed, uint64_t env_id0, double *sf, int32_t *si)
{
    const int nf = orc_nf(c), ni = orc_ni(c);
    for (int64_t b = 0; b < B; b++) orc_reset(c, seed, env_id0 + (uint64_t)b, sf + b * nf, si + b * ni);
}

void orc_step_batch(const orc_config *c, int64_t B, double *sf, int32_t *si, const int32_t *actions,
                    double *obs, double *reward, double *dist, int32_t *out_i,
                    int auto_reset, uint64_t seed, uint64_t env_id0)
{
    const int nf = orc_nf(c), ni = orc_ni(c), N = c->N, D = orc_obs_dim(c);
    for (int64_t b = 0; b < B; b++) {
        orc_step(c, sf + b * nf, si + b * ni, actions + b * N, obs + b * N * D, reward + b * N,
                 dist + b * N, out_i + b * 6);
        if (auto_reset && out_i[b * 6 + 1]) orc_reset(c, seed, env_id0 + (uint64_t)b, sf + b * nf, si + b * ni);
    }
}

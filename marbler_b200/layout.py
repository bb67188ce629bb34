"""Names of the structure-of-arrays state rows the kernels use (csrc/common.cuh) and the conversion
between that packed layout and the reference's own state variables (dict of arrays with a leading
env axis): poses (B,3,N), prev_pose (B,3,N), prev_valid, episode_steps, episode_count and per scenario
prey_loc/prey_sensed/prey_captured | loaded | load/zone_load/messages |
grid/goal_col/pixel_type/reached_goal | goal.  Host-side numpy, never on the step path.  The product converts
through the C ABI (mrb_get_state / mrb_set_state, csrc/state_io.cu); this module is the independent description
of the row layout that the tests check that conversion against."""
import numpy as np


def rows(scenario, N, P):
    f = 5 * N + 1 + {"PredatorCapturePrey": 2 * P, "Simple": 2}.get(scenario, 0)
    i = 3 + {"PredatorCapturePrey": 2, "Warehouse": 1, "MaterialTransport": N + 3, "ArcticTransport": 9}.get(scenario, 0)
    return f, i


def _mask(bits):                         # (B, K) 0/1 -> (B,) int32 bit mask
    bits = np.asarray(bits).astype(np.uint64)
    w = (bits << np.arange(bits.shape[1], dtype=np.uint64)[None, :]).sum(axis=1)
    return w.astype(np.uint32).view(np.int32)


def _unmask(word, K):
    w = np.asarray(word).view(np.uint32).astype(np.uint64)
    return ((w[:, None] >> np.arange(K, dtype=np.uint64)[None, :]) & 1).astype(np.uint8)


def _pack2(vals):                        # (B, K<=16) values in 0..3 -> (B,) int32, 2 bits each
    v = np.asarray(vals).astype(np.uint64) & 3
    w = (v << (2 * np.arange(v.shape[1], dtype=np.uint64))[None, :]).sum(axis=1)
    return w.astype(np.uint32).view(np.int32)


def _unpack2(word, K):
    w = np.asarray(word).view(np.uint32).astype(np.uint64)
    return ((w[:, None] >> (2 * np.arange(K, dtype=np.uint64))[None, :]) & 3).astype(np.int32)


def pack(scenario, N, P, st, B):
    """dict of (B, ...) arrays -> (state_f64 [rows, B], state_i32 [rows, B])."""
    nf, ni = rows(scenario, N, P)
    sf = np.zeros((nf, B), dtype=np.float64)
    si = np.zeros((ni, B), dtype=np.int32)

    def get(k, shape, dt):
        return np.asarray(st[k], dtype=dt).reshape((B,) + shape)
    sf[0:3 * N] = get("poses", (3 * N,), np.float64).T
    if "prev_pose" in st:
        sf[3 * N:5 * N] = get("prev_pose", (3, N), np.float64)[:, :2].reshape(B, 2 * N).T
    if "episode_return" in st:
        sf[5 * N] = get("episode_return", (), np.float64)
    si[0] = get("episode_steps", (), np.int32) if "episode_steps" in st else 0
    si[1] = get("prev_valid", (), np.int32) if "prev_valid" in st else 0
    si[2] = get("episode_count", (), np.int32) if "episode_count" in st else 0
    if scenario == "PredatorCapturePrey":
        sf[5 * N + 1:] = get("prey_loc", (2 * P,), np.float64).T
        si[3] = _mask(get("prey_sensed", (P,), np.int64))
        si[4] = _mask(get("prey_captured", (P,), np.int64))
    elif scenario == "Warehouse":
        si[3] = _mask(get("loaded", (N,), np.int64))
    elif scenario == "MaterialTransport":
        si[3:3 + N] = get("load", (N,), np.int32).T
        si[3 + N:5 + N] = get("zone_load", (2,), np.int32).T
        si[5 + N] = _pack2(get("messages", (4,), np.int64))
    elif scenario == "ArcticTransport":
        grid = get("grid", (96,), np.int64)
        for w in range(6):
            si[3 + w] = _pack2(grid[:, 16 * w:16 * w + 16])
        si[9] = get("goal_col", (), np.int32)
        si[10] = _pack2(get("pixel_type", (N,), np.int64))
        si[11] = _mask(get("reached_goal", (N,), np.int64))
    elif scenario == "Simple":
        sf[5 * N + 1:] = get("goal", (2,), np.float64).T
    return sf, si


def unpack(scenario, N, P, sf, si):
    """(state_f64 [rows, B], state_i32 [rows, B]) -> dict of (B, ...) arrays."""
    B = sf.shape[1]
    prev = np.zeros((B, 3, N))
    prev[:, :2] = sf[3 * N:5 * N].T.reshape(B, 2, N)
    st = {"poses": sf[0:3 * N].T.reshape(B, 3, N).copy(), "prev_pose": prev,
          "episode_return": sf[5 * N].copy(),
          "episode_steps": si[0].copy(), "prev_valid": si[1].copy(), "episode_count": si[2].copy()}
    if scenario == "PredatorCapturePrey":
        st["prey_loc"] = sf[5 * N + 1:].T.reshape(B, P, 2).copy()
        st["prey_sensed"] = _unmask(si[3], P)
        st["prey_captured"] = _unmask(si[4], P)
    elif scenario == "Warehouse":
        st["loaded"] = _unmask(si[3], N)
    elif scenario == "MaterialTransport":
        st["load"] = si[3:3 + N].T.copy()
        st["zone_load"] = si[3 + N:5 + N].T.copy()
        st["messages"] = _unpack2(si[5 + N], 4)
    elif scenario == "ArcticTransport":
        st["grid"] = np.concatenate([_unpack2(si[3 + w], 16) for w in range(6)], axis=1).reshape(B, 8, 12).astype(np.uint8)
        st["goal_col"] = si[9].copy()
        st["pixel_type"] = _unpack2(si[10], N)
        st["reached_goal"] = _unmask(si[11], N)
    elif scenario == "Simple":
        st["goal"] = sf[5 * N + 1:].T.copy()
    return st

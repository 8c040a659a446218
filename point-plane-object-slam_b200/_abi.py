"""ctypes mirror of include/ppo_ba.h and include/ppo_synth.h (structs only; no compute)."""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)

PPO_OK, PPO_E_INVALID, PPO_E_CUDA, PPO_E_NCCL, PPO_E_NOGPU, PPO_E_EMPTY = 0, -1, -2, -3, -4, -5
EDGE_POINT, EDGE_PLANE, EDGE_CUBOID_CAM, EDGE_POINT_CUBOID, EDGE_CUBOID_PLANE = range(5)
EDGE_KINDS = 5
PLANE_OBS, PLANE_VER, PLANE_PAR = 0, 1, 2
CUBOID_BBOX, CUBOID_CORNER, CUBOID_SE3 = 0, 1, 2
CU_FIXROLLPITCH, CU_FIXHEIGHT = 1, 2
EF_LEVEL1, EF_ROBUST = 1, 2
SOLVER_DENSE_X, SOLVER_6_3 = 0, 1
TRACE_MAX = 64

_pd = C.POINTER(C.c_double)
_pf = C.POINTER(C.c_float)
_pi = C.POINTER(C.c_int32)
_pb = C.POINTER(C.c_uint8)


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "huber_mono", "huber_stereo", "huber_plane", "huber_vp_plane", "huber_bbox", "huber_corner",
        "huber_cuboid_plane", "chi2_mono", "chi2_stereo", "chi2_plane", "chi2_vp_plane", "norm_bbox",
        "norm_corner", "lm_tau", "lm_good_upper", "lm_good_lower")] + [
        ("lm_max_trials", C.c_int32), ("solver", C.c_int32), ("iters_round1", C.c_int32),
        ("iters_round2", C.c_int32), ("ptcu_max_outside_margin_ratio", C.c_double),
        ("ptcu_prior_weight", C.c_double), ("huber_se3", C.c_double), ("norm_se3", C.c_double)]


class Graph(C.Structure):
    _fields_ = [
        ("n_kf", C.c_int32), ("kf_pose", _pd), ("kf_fixed", _pb), ("kf_intr", _pf),
        ("n_pt", C.c_int32), ("pt_xyz", _pd), ("pt_fixed", _pb),
        ("n_pl", C.c_int32), ("pl_coef", _pd),
        ("n_cu", C.c_int32), ("cu_state", _pd), ("cu_flags", _pb),
        ("pt_rowptr", _pi), ("n_pe", C.c_int32), ("pe_kf", _pi), ("pe_obs", _pf), ("pe_invsigma2", _pf),
        ("n_ple", C.c_int32), ("ple_plane", _pi), ("ple_kf", _pi), ("ple_kind", _pb), ("ple_meas", _pd),
        ("ple_info", _pd),
        ("n_cbe", C.c_int32), ("cbe_kf", _pi), ("cbe_cuboid", _pi), ("cbe_kind", _pb), ("cbe_meas", _pd),
        ("cbe_info", _pd),
        ("n_pce", C.c_int32), ("pce_cuboid", _pi), ("pce_rowptr", _pi), ("pce_pts", _pd),
        ("n_cpe", C.c_int32), ("cpe_cuboid", _pi), ("cpe_plane", _pi), ("cpe_meas", _pd), ("cpe_info", _pd),
    ]


class State(C.Structure):
    _fields_ = [("kf_pose", _pd), ("pt_xyz", _pd), ("pl_coef", _pd), ("cu_state", _pd)]


class Iter(C.Structure):
    _fields_ = [("chi2_before", C.c_double), ("chi2_after", C.c_double), ("lambda_", C.c_double),
                ("rho", C.c_double), ("trials", C.c_int32), ("accepted", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("terminated", C.c_int32), ("n_pose_dim", C.c_int32),
                ("n_landmarks", C.c_int32), ("n_active_edges", C.c_int32), ("total_trials", C.c_int32),
                ("chi2_initial", C.c_double), ("chi2_final", C.c_double), ("ms_total", C.c_double),
                ("ms_linearize", C.c_double), ("ms_schur", C.c_double), ("ms_solve", C.c_double),
                ("ms_update", C.c_double), ("trace", Iter * TRACE_MAX)]

    def trace_list(self):
        return [dict(chi2_before=t.chi2_before, chi2_after=t.chi2_after, lam=t.lambda_, rho=t.rho,
                     trials=t.trials, accepted=t.accepted)
                for t in self.trace[:min(self.iterations, TRACE_MAX)]]


class Result(C.Structure):
    _fields_ = [("round1", Stats), ("round2", Stats), ("n_outlier_point_edges", C.c_int32),
                ("n_outlier_plane_edges", C.c_int32), ("n_outlier_cuboid_edges", C.c_int32),
                ("skipped", C.c_int32)]


class SynthCfg(C.Structure):
    _fields_ = [("n_kf", C.c_int32), ("n_fixed", C.c_int32), ("n_pt", C.c_int32), ("n_pl", C.c_int32),
                ("n_cu", C.c_int32), ("seed", C.c_uint64), ("cuboid_2d", C.c_int32), ("corners_2d", C.c_int32),
                ("pt_obj_3d", C.c_int32), ("cuboid_plane", C.c_int32), ("plane_3d", C.c_int32),
                ("outlier_frac", C.c_double), ("stereo_frac", C.c_double), ("sort_points", C.c_int32),
                ("cuboid_3d", C.c_int32)]


_GRAPH_ARRAYS = {
    # field: (count expr, per-item, dtype)
    "kf_pose": ("n_kf", 7, np.float64), "kf_fixed": ("n_kf", 1, np.uint8), "kf_intr": ("n_kf", 5, np.float32),
    "pt_xyz": ("n_pt", 3, np.float64), "pt_fixed": ("n_pt", 1, np.uint8),
    "pl_coef": ("n_pl", 4, np.float64), "cu_state": ("n_cu", 10, np.float64), "cu_flags": ("n_cu", 1, np.uint8),
    "pt_rowptr": ("n_pt+1", 1, np.int32), "pe_kf": ("n_pe", 1, np.int32), "pe_obs": ("n_pe", 3, np.float32),
    "pe_invsigma2": ("n_pe", 1, np.float32),
    "ple_plane": ("n_ple", 1, np.int32), "ple_kf": ("n_ple", 1, np.int32), "ple_kind": ("n_ple", 1, np.uint8),
    "ple_meas": ("n_ple", 4, np.float64), "ple_info": ("n_ple", 3, np.float64),
    "cbe_kf": ("n_cbe", 1, np.int32), "cbe_cuboid": ("n_cbe", 1, np.int32), "cbe_kind": ("n_cbe", 1, np.uint8),
    "cbe_meas": ("n_cbe", 16, np.float64), "cbe_info": ("n_cbe", 1, np.float64),
    "pce_cuboid": ("n_pce", 1, np.int32), "pce_rowptr": ("n_pce+1", 1, np.int32), "pce_pts": ("n_pcp", 3, np.float64),
    "cpe_cuboid": ("n_cpe", 1, np.int32), "cpe_plane": ("n_cpe", 1, np.int32), "cpe_meas": ("n_cpe", 3, np.float64),
    "cpe_info": ("n_cpe", 3, np.float64),
}
_COUNTS = ("n_kf", "n_pt", "n_pl", "n_cu", "n_pe", "n_ple", "n_cbe", "n_pce", "n_cpe")


class GraphArrays:
    """numpy-owned flat graph; `.c` is the ppo_ba_graph view handed to the C-ABI / oracle."""

    def __init__(self, **arrays):
        self.a = {}
        for k, (_, per, dt) in _GRAPH_ARRAYS.items():
            v = arrays.get(k)
            if v is None:
                continue
            v = np.ascontiguousarray(v, dtype=dt)
            self.a[k] = v.reshape(-1, per) if per > 1 else v.reshape(-1)
        self._rebuild()

    def _rebuild(self):
        g = Graph()
        a = self.a
        g.n_kf = len(a["kf_pose"])
        g.n_pt = len(a["pt_xyz"]) if "pt_xyz" in a else 0
        g.n_pl = len(a["pl_coef"]) if "pl_coef" in a else 0
        g.n_cu = len(a["cu_state"]) if "cu_state" in a else 0
        g.n_pe = len(a["pe_kf"]) if "pe_kf" in a else 0
        g.n_ple = len(a["ple_plane"]) if "ple_plane" in a else 0
        g.n_cbe = len(a["cbe_kf"]) if "cbe_kf" in a else 0
        g.n_pce = len(a["pce_cuboid"]) if "pce_cuboid" in a else 0
        g.n_cpe = len(a["cpe_cuboid"]) if "cpe_cuboid" in a else 0
        if "pt_rowptr" not in a:
            a["pt_rowptr"] = np.zeros(g.n_pt + 1, np.int32)
        if "pce_rowptr" not in a:
            a["pce_rowptr"] = np.zeros(g.n_pce + 1, np.int32)
        for k, v in a.items():
            ct = {np.float64: _pd, np.float32: _pf, np.int32: _pi, np.uint8: _pb}[v.dtype.type]
            setattr(g, k, v.ctypes.data_as(ct))
        self.c = g

    @classmethod
    def from_c(cls, g):
        """Deep-copies a ppo_ba_graph (e.g. from ppo_synth_graph) into numpy arrays."""
        n = {k: getattr(g, k) for k in _COUNTS}
        n["n_pcp"] = 0
        if g.n_pce:
            n["n_pcp"] = int(np.ctypeslib.as_array(g.pce_rowptr, shape=(g.n_pce + 1,))[-1])
        arrays = {}
        for k, (cnt, per, dt) in _GRAPH_ARRAYS.items():
            p = getattr(g, k)
            if not p:
                continue
            count = eval(cnt, {}, n)
            if count <= 0:
                arrays[k] = np.zeros((0, per) if per > 1 else (0,), dt)
                continue
            arrays[k] = np.ctypeslib.as_array(p, shape=(count * per,)).copy()
        return cls(**arrays)

    def __getitem__(self, k):
        return self.a[k]

    def copy(self):
        return GraphArrays(**{k: v.copy() for k, v in self.a.items()})

    def nbytes(self):
        return int(sum(v.nbytes for v in self.a.values()))


class StateArrays:
    def __init__(self, g):
        self.kf_pose = np.zeros((g.n_kf, 7))
        self.pt_xyz = np.zeros((g.n_pt, 3))
        self.pl_coef = np.zeros((g.n_pl, 4))
        self.cu_state = np.zeros((g.n_cu, 10))
        s = State()
        s.kf_pose = self.kf_pose.ctypes.data_as(_pd)
        s.pt_xyz = self.pt_xyz.ctypes.data_as(_pd)
        s.pl_coef = self.pl_coef.ctypes.data_as(_pd)
        s.cu_state = self.cu_state.ctypes.data_as(_pd)
        self.c = s


def dptr(a):
    return a.ctypes.data_as(_pd)


def bptr(a):
    return a.ctypes.data_as(_pb)

"""ctypes binding of libdiffco_b200.so — the C ABI declared in include/diffco_b200.h.

The product path has no CPU fallback: if the shared object is missing the import of the native layer raises
with the build command, and every compute entry point raises if CUDA is unavailable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libdiffco_b200.so")

DC_MAX_DOF = 16
DC_MAX_LINKS = 16
DC_MAX_KEYPOINTS = 16
DC_MAX_ARMS = 2
DC_MAX_ARM_JOINTS = 8
DC_MAX_TOOL_POINTS = 2
DC_MAX_FEATURES = 64
DC_MAX_TREE_NODES = 24
DC_MAX_CLASSES = 8

ABI_VERSION = 7
DC_F32, DC_F64 = 0, 1
DC_FK_NONE, DC_FK_PLANAR_CHAIN, DC_FK_SE2_BODY, DC_FK_SE3_BODY, DC_FK_DH_ARMS, DC_FK_SE2_BASE_PLANAR_ARM, DC_FK_JOINT_TREE = range(7)
DC_JOINT_FIXED, DC_JOINT_REV_X, DC_JOINT_REV_Y, DC_JOINT_REV_Z, DC_JOINT_PRISMATIC = range(5)
DC_K_RQ, DC_K_POLYHARMONIC, DC_K_MULTIQUADRIC, DC_K_RQ_TEMPORAL = 1, 2, 3, 4
DC_GRAD_NONE, DC_GRAD_SUM, DC_GRAD_JAC = 0, 1, 2


class DhArm(C.Structure):
    _fields_ = [
        ("n_joints", C.c_int32),
        ("n_tool", C.c_int32),
        ("joint_index", C.c_int32 * DC_MAX_ARM_JOINTS),
        ("out_slot", C.c_int32 * DC_MAX_ARM_JOINTS),
        ("tool_slot", C.c_int32 * DC_MAX_TOOL_POINTS),
        ("a", C.c_double * DC_MAX_ARM_JOINTS),
        ("d", C.c_double * DC_MAX_ARM_JOINTS),
        ("s_alpha", C.c_double * DC_MAX_ARM_JOINTS),
        ("c_alpha", C.c_double * DC_MAX_ARM_JOINTS),
        ("theta0", C.c_double * DC_MAX_ARM_JOINTS),
        ("base", C.c_double * 12),
        ("offset", C.c_double * 3),
        ("tool", (C.c_double * 3) * DC_MAX_TOOL_POINTS),
    ]


class TreeNode(C.Structure):
    _fields_ = [
        ("parent", C.c_int32),
        ("q_index", C.c_int32),
        ("joint", C.c_int32),
        ("out_slot", C.c_int32),
        ("rot", C.c_double * 9),
        ("trans", C.c_double * 3),
        ("axis", C.c_double * 3),
        ("mimic_mul", C.c_double),
        ("mimic_off", C.c_double),
    ]


class FkDesc(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("dof", C.c_int32),
        ("n_points", C.c_int32),
        ("point_dim", C.c_int32),
        ("n_arms", C.c_int32),
        ("n_keypoints", C.c_int32),
        ("n_links", C.c_int32),
        ("n_repeat", C.c_int32),
        ("time_last", C.c_int32),
        ("n_nodes", C.c_int32),
        ("link_length", C.c_double * DC_MAX_LINKS),
        ("keypoints", (C.c_double * DC_MAX_KEYPOINTS) * 3),
        ("arms", DhArm * DC_MAX_ARMS),
        ("tree", TreeNode * DC_MAX_TREE_NODES),
    ]

    @property
    def n_features(self) -> int:
        return self.dof if self.type == DC_FK_NONE else self.n_points * self.point_dim


class KernelDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("order", C.c_int32), ("param", C.c_double), ("param2", C.c_double),
                ("alpha", C.c_double), ("order2", C.c_int32), ("reserved", C.c_int32)]


class Supports(C.Structure):
    _fields_ = [
        ("table", C.c_void_p),
        ("n", C.c_int64),
        ("n_features", C.c_int32),
        ("n_class", C.c_int32),
        ("f_pad", C.c_int32),
        ("row_stride", C.c_int32),
        ("dtype", C.c_int32),
        ("reserved", C.c_int32),
        ("tc_blob", C.c_void_p),
        ("tc_s2max", C.c_double),
        ("tc_gamma", C.c_double),
        ("table_lo", C.c_void_p),
    ]


DC_MAX_PEERS = 8


class PeerHandle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 64)]


class PeerTable(C.Structure):
    _fields_ = [("ptr", C.c_void_p * DC_MAX_PEERS)]


class TrajParams(C.Structure):
    _fields_ = [
        ("dif_weight", C.c_double), ("max_move_weight", C.c_double), ("collision_weight", C.c_double),
        ("joint_limit_weight", C.c_double), ("safety_bias", C.c_double), ("max_speed", C.c_double),
        ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
        ("limits", (C.c_double * 2) * DC_MAX_DOF),
        ("wrap", C.c_int32 * DC_MAX_DOF),
    ]


class TrajDense(C.Structure):
    _fields_ = [("count", C.c_void_p), ("seg_offset", C.c_void_p), ("score", C.c_void_p), ("score_grad", C.c_void_p),
                ("max_points", C.c_int32), ("reserved", C.c_int32)]


DC_OPT_TC_ENABLE, DC_OPT_TC_ERR_COEF, DC_OPT_TC_TOL_PAIR, DC_OPT_TC_MIN_BATCH, DC_OPT_TC_STATS, DC_OPT_PEER_TIMEOUT_S = 1, 2, 3, 4, 5, 6
KERNEL_NAMES = {0: "lane-split", 1: "thread-per-query", 2: "tensor-core", -1: "none"}


# name -> (restype, argtypes); must list every function declared in include/diffco_b200.h
PROTOTYPES = {
    "dc_abi_version": (C.c_int, []),
    "dc_status_string": (C.c_char_p, [C.c_int]),
    "dc_launch_count": (C.c_int64, []),
    "dc_supports_layout": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "dc_pack_supports": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "dc_supports_tc_bytes": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
    "dc_pack_supports_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(KernelDesc), C.c_void_p, C.c_void_p]),
    "dc_supports_tc_info": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "dc_set_option": (C.c_int, [C.c_int32, C.c_double]),
    "dc_get_option": (C.c_double, [C.c_int32]),
    "dc_last_score_kernel": (C.c_int, []),
    "dc_score_grad": (C.c_int, [C.POINTER(FkDesc), C.POINTER(KernelDesc), C.POINTER(Supports), C.c_void_p, C.c_int64,
                                C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "dc_host_pipeline_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64, C.c_int32]),
    "dc_host_pipeline_destroy": (None, [C.c_void_p]),
    "dc_score_grad_host": (C.c_int, [C.c_void_p, C.POINTER(FkDesc), C.POINTER(KernelDesc), C.POINTER(Supports), C.c_void_p,
                                     C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "dc_kernel_matrix": (C.c_int, [C.POINTER(KernelDesc), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                   C.c_int32, C.c_void_p, C.c_void_p]),
    "dc_fk_forward": (C.c_int, [C.POINTER(FkDesc), C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "dc_fk_forward_split": (C.c_int, [C.POINTER(FkDesc), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dc_pack_supports_lo": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "dc_perceptron_train": (C.c_int, [C.POINTER(KernelDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                      C.c_void_p, C.c_void_p]),
    "dc_perceptron_train_rows": (C.c_int, [C.POINTER(KernelDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                           C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "dc_peer_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.POINTER(PeerHandle)]),
    "dc_peer_open": (C.c_int, [C.POINTER(PeerHandle), C.POINTER(C.c_void_p)]),
    "dc_peer_close": (C.c_int, [C.c_void_p]),
    "dc_peer_free": (C.c_int, [C.c_void_p]),
    "dc_peer_barrier": (C.c_int, [C.POINTER(PeerTable), C.c_int32, C.c_int32, C.c_uint32, C.c_void_p]),
    "dc_score_grad_bcast": (C.c_int, [C.POINTER(FkDesc), C.POINTER(KernelDesc), C.POINTER(Supports), C.c_void_p, C.c_int64,
                                      C.POINTER(PeerTable), C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "dc_score_grad_bcast_sync": (C.c_int, [C.POINTER(FkDesc), C.POINTER(KernelDesc), C.POINTER(Supports), C.c_void_p, C.c_int64,
                                      C.POINTER(PeerTable), C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.POINTER(PeerTable), C.c_int32, C.c_uint32, C.c_void_p]),
    "dc_traj_step": (C.c_int, [C.POINTER(FkDesc), C.POINTER(TrajParams), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dc_traj_dense_path": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "dc_traj_step_ex": (C.c_int, [C.POINTER(FkDesc), C.POINTER(TrajParams), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.POINTER(TrajDense), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_double, C.c_void_p]),
    "dc_fk_vjp": (C.c_int, [C.POINTER(FkDesc), C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dc_fk_tree_frames": (C.c_int, [C.POINTER(FkDesc), C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load():
    """Load libdiffco_b200.so (once).  Raises NativeLibraryError with the build recipe if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m diffco_b200.build` (nvcc, sm_100a). "
            "diffco_b200 has no CPU or PyTorch fallback for the collision-score path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here == ABI mismatch, let it propagate
        fn.restype = res
        fn.argtypes = args
    if lib.dc_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"ABI version mismatch: library reports {lib.dc_abi_version()}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().dc_status_string(status).decode()
        raise RuntimeError(f"{what} failed: {msg} (dc_status {status})")

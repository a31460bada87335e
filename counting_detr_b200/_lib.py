"""ctypes loader for libcdetr_sm100a.so (the C ABI declared in include/cdetr.h).

The product path has no fallback: if the library is missing or a call fails, this raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CDETR_LIB_PATH") or os.path.join(_HERE, "lib", "libcdetr_sm100a.so")   # override: A/B builds
_lib = None


class CdetrError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CdetrError(
                f"{LIB_PATH} not built: run `python __graft_entry__.py` (nvcc, sm_100a). "
                "There is no CPU/eager fallback for the hot path.")
        _lib = C.CDLL(LIB_PATH)
        _lib.cdetr_last_error.restype = C.c_char_p
        _lib.cdetr_version.restype = C.c_int
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise CdetrError(f"{what} failed ({rc}): {lib().cdetr_last_error().decode()}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class SplitT(C.Structure):
    """cdetr_split_t"""
    _fields_ = [("base", C.c_void_p), ("ld", C.c_int64), ("plane", C.c_int64)]


def split_view(t):
    """t: bf16 tensor [2, rows, ld] (planes hi, lo), possibly a row/column slice of a bigger one."""
    if t is None:
        return SplitT(None, 0, 0)
    assert t.dtype == torch.bfloat16 and t.dim() == 3 and t.shape[0] == 2 and t.stride(2) == 1, (t.shape, t.stride())
    return SplitT(t.data_ptr(), t.stride(1), t.stride(0))


class PackEntryT(C.Structure):
    """cdetr_pack_entry_t"""
    _fields_ = [("w", C.c_void_p), ("bn_w", C.c_void_p), ("bn_b", C.c_void_p), ("bn_rm", C.c_void_p), ("bn_rv", C.c_void_p),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("dst", SplitT), ("dst_t", SplitT), ("dst_d", SplitT),
                ("cout", C.c_int32), ("cin", C.c_int32), ("taps", C.c_int32), ("pad_", C.c_int32)]


class GemmT(C.Structure):
    """cdetr_gemm_t"""
    _fields_ = [
        ("mode", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a", SplitT), ("b", SplitT),
        ("block_n", C.c_int32), ("split_k", C.c_int32),
        ("row_scale", C.c_void_p), ("bias", C.c_void_p),
        ("add_split", SplitT),
        ("add_f32", C.c_void_p), ("ld_add_f32", C.c_int64),
        ("mask", SplitT),
        ("relu", C.c_int32), ("accumulate", C.c_int32),
        ("out_f32", C.c_void_p), ("ld_out_f32", C.c_int64),
        ("out_split", SplitT),
        ("conv_taps", C.c_int32), ("conv_H", C.c_int32), ("conv_W", C.c_int32), ("conv_C", C.c_int32),
        ("conv_dil", C.c_int32), ("conv_sign", C.c_int32),
        ("pass_mask", C.c_int32),
    ]


def _ptr(t):
    return None if t is None else t.data_ptr()


# kernel-launch accounting (bench.py's gpu_launches) and optional per-GEMM event trace (bench.py's roofline)
COUNTER = {"launches": 0}
_LAUNCHES = {"cdetr_exemplar_concat": 2, "cdetr_exemplar_concat_bwd": 2, "cdetr_rcda_bwd": 3, "cdetr_rcda_bwd_kv": 2, "cdetr_mha_bwd": 2,
             "cdetr_mt_grad_norm": 2, "cdetr_mt_adamw": 2}
GEMM_TRACE = None
SKIP = None            # tools/criticality.py only: set of entry points NOT launched (timing experiments, wrong results)
CALL_TRACE = None      # when a list: (name, leading int args, start event, end event) of every attention-core call


def gemm(a, b, M, N, K, mode=0, out_f32=None, out_split=None, bias=None, row_scale=None,
         add_split=None, add_f32=None, mask=None, relu=False, accumulate=False, block_n=0, split_k=1, conv=None,
         pass_mask=0):
    """cdetr_gemm: see include/cdetr.h. a, b, add_split, mask, out_split are split tensors [2, rows, ld].
    conv = (H, W, C, dil, sign) turns the conv operand (a in mode 0, b in mode 1) into an implicit 3x3 im2col."""
    g = GemmT()
    if conv is not None:
        g.conv_taps = 9
        g.conv_H, g.conv_W, g.conv_C, g.conv_dil, g.conv_sign = conv
    g.mode, g.M, g.N, g.K = mode, M, N, K
    g.a, g.b = split_view(a), split_view(b)
    g.block_n, g.split_k = block_n, split_k
    g.pass_mask = pass_mask
    g.row_scale, g.bias = _ptr(row_scale), _ptr(bias)
    g.add_split = split_view(add_split)
    g.add_f32 = _ptr(add_f32)
    g.ld_add_f32 = add_f32.stride(0) if add_f32 is not None else 0
    g.mask = split_view(mask)
    g.relu, g.accumulate = int(relu), int(accumulate)
    g.out_f32 = _ptr(out_f32)
    g.ld_out_f32 = out_f32.stride(0) if out_f32 is not None else 0
    g.out_split = split_view(out_split)
    COUNTER["launches"] += 1
    if GEMM_TRACE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib().cdetr_gemm(C.byref(g), stream_ptr()), "cdetr_gemm")
        e1.record()
        GEMM_TRACE.append((M, N, K, e0, e1))
        return
    check(lib().cdetr_gemm(C.byref(g), stream_ptr()), "cdetr_gemm")


def to_split(x):
    """torch restatement of the split-bf16 format (used by tests; the product path packs on device)."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo])


def from_split(s):
    return s[0].float() + s[1].float()


# ---------------------------------------------------------------------------------------------
# Generic typed call layer for the rest of the C ABI (include/cdetr.h).
#   p = device pointer (torch tensor or None), i = int32, l = int64, f = float, S = cdetr_split_t
# ---------------------------------------------------------------------------------------------
_SIGS = {
    "cdetr_bn_fold": "ppppfipp",
    "cdetr_pack_weight": "piiipSS",
    "cdetr_pack_weight_dgrad": "piiipS",
    "cdetr_unpack_conv_grad": "piiip",
    "cdetr_mt_pack_weights": "ppiif",
    "cdetr_to_split": "plilS",
    "cdetr_from_split": "Slipl",
    "cdetr_stem_im2col": "piiiS",
    "cdetr_stem_conv": "piiiSpS",
    "cdetr_im2col3x3": "SiiiiiiS",
    "cdetr_col2im3x3": "SiiiiiiSS",
    "cdetr_maxpool3x3s2": "SiiiiS",
    "cdetr_subsample2": "SiiiiS",
    "cdetr_upsample2_zero": "SiiiiS",
    "cdetr_exemplar_concat": "SiiiipipS",
    "cdetr_exemplar_concat_bwd": "SSpiiiipipSS",
    "cdetr_groupnorm_fwd": "piiiippfpSp",
    "cdetr_groupnorm_bwd": "ppiiiipppSpp",
    "cdetr_layernorm_fwd": "pplippfppSp",
    "cdetr_layernorm_bwd": "pppplippSpp",
    "cdetr_sine_embed": "pliiiip",
    "cdetr_sine_embed_bwd": "pliiiipp",
    "cdetr_add_bcast": "ppliiiilpS",
    "cdetr_reduce_axis": "piiiiifpipS",
    "cdetr_combine_bcast": "ppppfpfliiip",
    "cdetr_colsum": "pSllip",
    "cdetr_box_head_fwd": "pplp",
    "cdetr_box_head_bwd": "ppplpSp",
    "cdetr_scale": "plf",
    "cdetr_mask_prepare": "piiiiipppp",
    "cdetr_exemplar_centres": "piiipp",
    "cdetr_rcda_fwd": "iiiiiipppppppppS",
    "cdetr_rcda_fwd_tc": "iiiiiippppSppppS",
    "cdetr_rcda_bwd_q_tc": "iiiiiippSpppppSS",
    "cdetr_rcda_bwd_v_tc": "iiiiiippSS",
    "cdetr_rcda_bwd_k": "iiiiiippppSS",
    "cdetr_rcda_bwd_kv": "iiiiiipppppppSSS",
    "cdetr_rcda_bwd": "iiiiiippppppppppSSSSS",
    "cdetr_mha_fwd": "iiiippplSp",
    "cdetr_mha_bwd": "iiiippplSpppSSS",
    "cdetr_match_cost": "pipppiiifffp",
    "cdetr_lsap": "ppiiipppp",
    "cdetr_set_loss_fwd": "ppppppppiiipffpppppppp",
    "cdetr_set_loss_bwd": "pppppplppp",
    "cdetr_bbox_loss_fwd": "ppplppp",
    "cdetr_bbox_loss_bwd": "ppplp",
    "cdetr_postprocess_topk": "pppiiiippp",
    "cdetr_infer_select": "pipppiifpppppp",
    "cdetr_pseudo_label_format": "ppplpp",
    "cdetr_normalize_u8": "piiiHHp",
    "cdetr_mt_grad_norm": "ppiifpp",
    "cdetr_mt_clip_scale": "ppiip",
    "cdetr_mt_adamw": "ppiipfffpp",
}
_CT = {"p": C.c_void_p, "i": C.c_int32, "l": C.c_int64, "f": C.c_float, "S": SplitT,
       "H": C.c_void_p}    # H = HOST pointer to a small float array (ctypes array)
_bound = {}


def _bind(name):
    fn = _bound.get(name)
    if fn is None:
        fn = getattr(lib(), name)
        fn.argtypes = [_CT[c] for c in _SIGS[name]] + [C.c_void_p]
        fn.restype = C.c_int
        _bound[name] = fn
    return fn


def call(name, *args):
    """Invoke a C-ABI entry point on torch's current stream; tensors become raw device pointers."""
    sig = _SIGS[name]
    assert len(args) == len(sig), (name, len(args), len(sig))
    conv = []
    for a, c in zip(args, sig):
        if c == "p":
            conv.append(None if a is None else (a.data_ptr() if isinstance(a, torch.Tensor) else a))
        elif c == "S":
            conv.append(a if isinstance(a, SplitT) else split_view(a))
        elif c == "H":
            conv.append(C.cast(a, C.c_void_p))
        elif c == "f":
            conv.append(float(a))
        else:
            conv.append(int(a))
    if SKIP is not None and name in SKIP:
        return
    COUNTER["launches"] += _LAUNCHES.get(name, 1)
    if CALL_TRACE is not None and name.startswith(("cdetr_rcda", "cdetr_mha")):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_bind(name)(*conv, torch.cuda.current_stream().cuda_stream), name)
        e1.record()
        CALL_TRACE.append((name, tuple(conv[:6]), e0, e1))
        return
    check(_bind(name)(*conv, torch.cuda.current_stream().cuda_stream), name)


EXPORTED = ["cdetr_version", "cdetr_last_error", "cdetr_gemm", "cdetr_gemm_debug_timeline"] + list(_SIGS)

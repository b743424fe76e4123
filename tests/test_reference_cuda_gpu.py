"""Parity against the REFERENCE'S OWN CUDA kernels, on the B200, on the same device buffers.

north_star: "The reference is the dev/*.py PyTorch ground-truth layers and the reference CUDA kernels."
oracle/Makefile compiles the reference's dev/*.cu, from where they lie, into oracle/_ref/libunetcu_ref.so (-D LINKING
strips the main()s, as dev/Makefile:26 does) and the reference's own test programs, UNCHANGED, against
include/legacy/*.cuh + libunet_b200.so into oracle/_ref/legacy/*_legacy.  Both travel to the GPU box (git-ignored, not
gpurun-ignored); /root/reference itself is never read here.

  1. conv2d_k3_forward3 / conv2d_k3_backward2 at the reference's own benchmark shape B32 192->64 @64x64
     (dev/conv2d_k3.cu:2586-2590): reference kernels vs ub_* on identical inputs, bf16 tensor path (stated bound)
     and fp32 validation mode (1e-3 relative, north_star).
  2. resblock_forward / resblock_backward 192->64 (dev/resblock.cu:542-630 checks every intermediate): reference
     composites over reference kernels vs ub_resblock_* -- every intermediate activation and every gradient.
  3. The reference's test program dev/resblock.cu, compiled unchanged against the legacy headers, run against a fixture
     in its own file format (written here with the oracle, dev/resblock.py:318-358): in fp32 validation mode it passes
     its own 1e-5 / 1e-4 absolute checks ("All results match").
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from gpu_util import TOL_BF16, TOL_F32, rel_inf

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libunetcu_ref.so")
LEGACY = os.path.join(ROOT, "oracle", "_ref", "legacy")


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def ref():
    """ctypes handle of the reference library + a lookup of its C++ (mangled) entry points by demangled name."""
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libunetcu_ref.so not built (needs /root/reference at build time)")
    lib = C.CDLL(REF_SO)
    mangled = subprocess.run(["nm", "-D", "--defined-only", REF_SO], stdout=subprocess.PIPE, text=True).stdout.split("\n")
    plain = subprocess.run(["nm", "-D", "--defined-only", "-C", REF_SO], stdout=subprocess.PIPE, text=True).stdout.split("\n")
    table = {}
    for m, p in zip(mangled, plain):
        if " T " in m:
            table.setdefault(p.split(" T ", 1)[1].split("(")[0], m.split(" T ", 1)[1])

    def get(name):
        f = getattr(lib, table[name])
        f.restype = None
        return f

    handle = C.c_void_p()
    cublas = C.CDLL("libcublas.so.12")
    assert cublas.cublasCreate_v2(C.byref(handle)) == 0
    return get, handle


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def I(*v):
    return [C.c_int(int(x)) for x in v]


@pytest.fixture
def precision(ub):
    """Restores the bf16 default after a test that switches the layer API to fp32."""
    yield lambda mode: ub.lib().ub_set_layer_precision(mode)
    ub.lib().ub_set_layer_precision(0)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_conv3x3_vs_reference_kernels_at_the_reference_shape(ub, ref, precision, mode):
    get, _ = ref
    precision(1 if mode == "fp32" else 0)
    B, Ci, Co, H, W = 32, 192, 64, 64, 64  # dev/conv2d_k3.cu:2586-2590
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(B, Ci, H, W, device="cuda", generator=g)
    w = torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Ci)
    b = torch.randn(Co, device="cuda", generator=g) * 0.1
    dout = torch.randn(B, Co, H, W, device="cuda", generator=g) / math.sqrt(H * W)
    out_r, out_u = torch.zeros(B, Co, H, W, device="cuda"), torch.zeros(B, Co, H, W, device="cuda")
    get("conv2d_k3_forward3")(P(x), P(w), P(b), P(out_r), *I(B, Ci, Co, H, W))
    L = ub.lib()
    assert L.ub_conv2d_k3_forward3(P(x), P(w), P(b), P(out_u), *I(B, Ci, Co, H, W)) == 0, L.ub_last_error()
    torch.cuda.synchronize()
    tol_inf, tol_l2 = (TOL_F32, 1e-4) if mode == "fp32" else (TOL_BF16, 6e-3)
    assert rel_inf(out_u, out_r) < tol_inf and rel_l2(out_u, out_r) < tol_l2, (rel_inf(out_u, out_r), rel_l2(out_u, out_r))

    # backward: the reference needs its split-K scratch (dev/resblock.cu:225-226 sizes), ours ignores it
    wbuf = torch.zeros(Co * Ci * 9 * B * 32, device="cuda")  # dweight_buf as dev/resblock.cu:225 sizes it
    bbuf = torch.zeros(Co * B * 32, device="cuda")
    dx_r, dw_r, db_r = torch.zeros_like(x), torch.zeros_like(w), torch.zeros_like(b)
    dx_u, dw_u, db_u = torch.zeros_like(x), torch.zeros_like(w), torch.zeros_like(b)
    get("conv2d_k3_backward2")(P(dout), P(x), P(w), P(wbuf), P(bbuf), P(dx_r), P(dw_r), P(db_r), *I(B, Ci, Co, H, W))
    assert L.ub_conv2d_k3_backward2(P(dout), P(x), P(w), None, None, P(dx_u), P(dw_u), P(db_u), *I(B, Ci, Co, H, W)) == 0
    torch.cuda.synchronize()
    for name, u, r in (("dx", dx_u, dx_r), ("dweight", dw_u, dw_r), ("dbias", db_u, db_r)):
        assert rel_inf(u, r) < tol_inf and rel_l2(u, r) < tol_l2, (name, rel_inf(u, r), rel_l2(u, r))


RES_P = "gn1_w gn1_b cv3_1_w cv3_1_b l_emb_w l_emb_b gn2_w gn2_b cv3_2_w cv3_2_b res_cv1_w res_cv1_b".split()
RES_A = ("gn1 gn1_mean gn1_rstd silu1 ud_h ud_x cv3_1 silu_emb l_emb broad_emb add1 gn2 gn2_mean gn2_rstd silu2 cv3_2 "
         "res_cv1 add2").split()
RES_K = "buf_BCemb buf_BCHoWo buf1_BCHW buf2_BCHW dout dweight_buf dbias_buf".split()


class RefResParams(C.Structure):  # dev/resblock.cuh:8-24
    _fields_ = [(n, C.c_void_p) for n in RES_P] + [("param_sizes", C.c_size_t * 12), ("n_params", C.c_size_t)]


class RefResActs(C.Structure):  # dev/resblock.cuh:26-49
    _fields_ = ([(n, C.c_void_p) for n in RES_A] + [("act_sizes", C.c_size_t * 18), ("n_acts", C.c_size_t),
                                                    ("input", C.c_void_p), ("emb", C.c_void_p)])


class RefResBack(C.Structure):  # dev/resblock.cuh:55-67
    _fields_ = ([(n, C.c_void_p) for n in RES_K] + [("dx", C.c_void_p), ("demb", C.c_void_p),
                                                    ("back_sizes", C.c_size_t * 7), ("n_backs", C.c_size_t)])


def _ptr_struct(names):
    class S(C.Structure):
        _fields_ = [(n, C.c_void_p) for n in names]
    return S


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_resblock_vs_reference_composites(ub, ref, precision, mode):
    """dev/resblock.cu:542-630 compares every intermediate; here the comparison partner is the reference's own CUDA
    implementation of the block run on the same inputs."""
    get, cublas = ref
    precision(1 if mode == "fp32" else 0)
    B, Cc, Cemb, Co, H, W, G = 8, 192, 256, 64, 64, 64, 32
    g = torch.Generator(device="cuda").manual_seed(5)
    rn = lambda *s, sc=1.0: torch.randn(*s, device="cuda", generator=g) * sc
    shapes = {"gn1_w": (Cc,), "gn1_b": (Cc,), "cv3_1_w": (Co, Cc, 3, 3), "cv3_1_b": (Co,), "l_emb_w": (Co, Cemb),
              "l_emb_b": (Co,), "gn2_w": (Co,), "gn2_b": (Co,), "cv3_2_w": (Co, Co, 3, 3), "cv3_2_b": (Co,),
              "res_cv1_w": (Co, Cc), "res_cv1_b": (Co,)}
    scale = {"cv3_1_w": 1 / math.sqrt(9 * Cc), "cv3_2_w": 1 / math.sqrt(9 * Co), "l_emb_w": 1 / math.sqrt(Cemb),
             "res_cv1_w": 1 / math.sqrt(Cc)}
    params = {n: rn(*s, sc=scale.get(n, 0.3)) + (1.0 if n in ("gn1_w", "gn2_w") else 0.0) for n, s in shapes.items()}
    x, emb = rn(B, Cc, H, W), rn(B, Cemb)
    dout = rn(B, Co, H, W, sc=1 / (B * math.sqrt(H * W)))
    n_in, n_out, ng = B * Cc * H * W, B * Co * H * W, B * G
    act_sizes = [n_in, ng, ng, n_in, n_in, n_in, n_out, B * Cemb, B * Co, n_out, n_out, n_out, ng, ng, n_out, n_out,
                 n_out, n_out]
    back_sizes = [B * Cemb, n_in, n_in, n_in, n_out, Co * Cc * 9 * B * 32, Co * B * 32]

    def run(which):
        acts = {n: torch.zeros(s, device="cuda") for n, s in zip(RES_A, act_sizes)}
        grads = {n: torch.zeros_like(t) for n, t in params.items()}
        backs = {n: torch.zeros(s, device="cuda") for n, s in zip(RES_K, back_sizes)}
        xin, ein = x.clone(), emb.clone()  # backward overwrites input / emb with dx / demb (dev/resblock.cu:581-582)
        backs["dout"].copy_(dout.flatten())
        if which == "ref":
            rp, rg, ra, rk = RefResParams(), RefResParams(), RefResActs(), RefResBack()
            for n in RES_P:
                setattr(rp, n, params[n].data_ptr()), setattr(rg, n, grads[n].data_ptr())
            for n in RES_A:
                setattr(ra, n, acts[n].data_ptr())
            ra.input, ra.emb = xin.data_ptr(), ein.data_ptr()
            for n in RES_K:
                setattr(rk, n, backs[n].data_ptr())
            rk.dx, rk.demb = xin.data_ptr(), ein.data_ptr()
            args = [cublas] + I(Cc, Cemb, Co, B, H, W, 512, 0, 0, G)
            get("resblock_forward")(*args, C.byref(rp), C.byref(ra))
            torch.cuda.synchronize()
            fwd = {n: acts[n].clone() for n in RES_A}
            get("resblock_backward")(*args, C.byref(rp), C.byref(rg), C.byref(ra), C.byref(rk))
        else:
            UP, UA = _ptr_struct(RES_P), _ptr_struct(RES_A + ["input", "emb"])
            UK = _ptr_struct("buf_BCemb buf_BCHoWo buf1_BCHW buf2_BCHW dout dx demb".split())
            up_, ug, ua, uk = UP(), UP(), UA(), UK()
            for n in RES_P:
                setattr(up_, n, params[n].data_ptr()), setattr(ug, n, grads[n].data_ptr())
            for n in RES_A:
                setattr(ua, n, acts[n].data_ptr())
            ua.input, ua.emb = xin.data_ptr(), ein.data_ptr()
            for n in "buf_BCemb buf_BCHoWo buf1_BCHW buf2_BCHW dout".split():
                setattr(uk, n, backs[n].data_ptr())
            uk.dx, uk.demb = xin.data_ptr(), ein.data_ptr()
            L = ub.lib()
            args = I(Cc, Cemb, Co, B, H, W, 0, 0, G)
            assert L.ub_resblock_forward(*args, C.byref(up_), C.byref(ua)) == 0, L.ub_last_error()
            torch.cuda.synchronize()
            fwd = {n: acts[n].clone() for n in RES_A}
            assert L.ub_resblock_backward(*args, C.byref(up_), C.byref(ug), C.byref(ua), C.byref(uk)) == 0, L.ub_last_error()
        torch.cuda.synchronize()
        return fwd, grads, xin, ein

    fr, gr, dxr, der = run("ref")
    fu, gu, dxu, deu = run("ours")
    tol_inf, tol_l2 = (TOL_F32, 2e-4) if mode == "fp32" else (TOL_BF16, 1e-2)
    worst = []
    for n in RES_A:  # every forward intermediate the reference materialises (ud_h / ud_x are aliases without up/down)
        if n in ("ud_h", "ud_x"):
            continue
        worst.append((rel_l2(fu[n], fr[n]), rel_inf(fu[n], fr[n]), "act " + n))
    for n in RES_P:
        worst.append((rel_l2(gu[n], gr[n]), rel_inf(gu[n], gr[n]), "grad " + n))
    worst.append((rel_l2(dxu, dxr), rel_inf(dxu, dxr), "dx"))
    worst.append((rel_l2(deu, der), rel_inf(deu, der), "demb"))
    worst.sort(reverse=True)
    print("worst tensors (rel-L2, rel-max):", worst[:5])
    for l2, inf, n in worst:
        assert l2 < tol_l2 and inf < tol_inf, (n, l2, inf, worst[:5])


def _write_resblock_fixture(d, oracle, B, Cc, Cemb, Co, H, W):
    """resblock_params.bin / resblock_states.bin in the reference's own format (dev/resblock.py:318-358), computed
    with the oracle's functional ops and autograd."""
    g = torch.Generator().manual_seed(7)
    rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc)
    names = RES_P if Cc != Co else RES_P[:10]
    shapes = {"gn1_w": (Cc,), "gn1_b": (Cc,), "cv3_1_w": (Co, Cc, 3, 3), "cv3_1_b": (Co,), "l_emb_w": (Co, Cemb),
              "l_emb_b": (Co,), "gn2_w": (Co,), "gn2_b": (Co,), "cv3_2_w": (Co, Co, 3, 3), "cv3_2_b": (Co,),
              "res_cv1_w": (Co, Cc), "res_cv1_b": (Co,)}
    sc = {"cv3_1_w": 1 / math.sqrt(9 * Cc), "cv3_2_w": 1 / math.sqrt(9 * Co), "l_emb_w": 1 / math.sqrt(Cemb),
          "res_cv1_w": 1 / math.sqrt(Cc)}
    Pm = {n: (rn(*shapes[n], sc=sc.get(n, 0.2)) + (1.0 if n in ("gn1_w", "gn2_w") else 0.0)).requires_grad_(True)
          for n in names}
    x = (rn(B, Cc, H, W) * (H * W) ** -0.5).requires_grad_(True)
    emb = (rn(B, Cemb) * Cemb ** -0.5).requires_grad_(True)
    h_gn1 = oracle.groupnorm(x, Pm["gn1_w"], Pm["gn1_b"])
    h_silu1 = oracle.silu(h_gn1)
    h_1 = oracle.conv3x3(h_silu1, Pm["cv3_1_w"], Pm["cv3_1_b"])
    x_1 = x
    emb_1 = torch.nn.functional.linear(oracle.silu(emb), Pm["l_emb_w"], Pm["l_emb_b"])
    emb_broad = emb_1[:, :, None, None].expand(B, Co, H, W).contiguous()
    h_plus = h_1 + emb_broad
    h_gn2 = oracle.groupnorm(h_plus, Pm["gn2_w"], Pm["gn2_b"])
    h_silu2 = oracle.silu(h_gn2)
    h_2 = oracle.conv3x3(h_silu2, Pm["cv3_2_w"], Pm["cv3_2_b"])
    skip = x_1 if Cc == Co else oracle.conv1x1(x_1, Pm["res_cv1_w"][:, :, None, None], Pm["res_cv1_b"])
    out = h_2 + skip
    dout = rn(B, Co, H, W) / (B * (H * W) ** 0.5)
    (out * dout).sum().backward()
    header = np.zeros(256, dtype=np.int32)
    header[:10] = [12345678, B, Cc, Cemb, Co, H, W, 0, 0, 32]
    w = lambda f, t: f.write(t.detach().contiguous().numpy().astype(np.float32).tobytes())
    with open(os.path.join(d, "resblock_params.bin"), "wb") as f:
        f.write(header.tobytes())
        for n in names:
            w(f, Pm[n])
    with open(os.path.join(d, "resblock_states.bin"), "wb") as f:
        for t in (x, emb, h_gn1, h_silu1, h_1, x_1, emb_1, h_plus, h_gn2, h_silu2, h_2, out, dout, x.grad, emb.grad,
                  emb_broad):
            w(f, t)
        for n in names:
            w(f, Pm[n].grad)


def test_legacy_programs_run(ub, oracle, tmp_path):
    """The reference's dev/resblock.cu test program -- compiled UNCHANGED against include/legacy/*.cuh, its own
    resblock_forward / resblock_backward composites calling THIS library's layer launchers -- passes its own checks
    (validate_result at 1e-5 forward / 1e-4 backward, dev/resblock.cu:542-630) in the fp32 validation mode."""
    exe = os.path.join(LEGACY, "resblock_legacy")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/legacy/resblock_legacy not built (needs /root/reference at build time)")
    _write_resblock_fixture(str(tmp_path), oracle, B=4, Cc=64, Cemb=256, Co=128, H=16, W=16)
    env = dict(os.environ, UB_LAYER_PRECISION="fp32",
               LD_LIBRARY_PATH=os.path.dirname(ub.LIB_PATH) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=600)
    tail = r.stdout[-1500:]
    assert r.returncode == 0, tail
    assert "Forward pass successful" in r.stdout and "All results match" in r.stdout, tail
    assert "Mismatch" not in r.stdout, tail

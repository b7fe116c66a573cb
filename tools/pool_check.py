"""Stage-level check of the image pool kernels (pass B of S9) against a float64 torch evaluation of the same algebra from
the kernel's own inputs (w_eff planes, cterm, xbar in the workspace): scaled scores, probabilities, weighted sums.
With a third argument it also times the BACK stage at bench size over a sweep of the producer's L2-prefetch distance (PT_POOL_PF).
Usage (GPU box): python tools/pool_check.py [views_per_scene] [scenes] [time | comma-separated PT_POOL_PF values] [mma | umma]
(`umma` / `mma` restrict the check to the tcgen05 or the mma.sync pool kernel, default both)"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PT_POOL_DEBUG", "64")
from proxytransformation_b200 import ProxyTransformationNormReverse, ops, _lib, synthetic as syn

V = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
C, HW, EMB, HEADS, TP, YA, WPITCH, SP = 512, 225, 256, 8, 228, 768, 528, 232
WPLANE, SBUF = HEADS * WPITCH, HEADS * SP
DBG = HEADS * 256
cfg = syn.C2_WIDE.replace(n_views=V)
m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
m.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True))
m = m.cuda()
torch.manual_seed(1)
img = (torch.relu(torch.randn(B, V, C, 15, 15, device="cuda")) * 1.5).bfloat16()
BV = B * V
need = _lib.load().pt_img_attnpool_ws_bytes(BV, C, HW, EMB, HEADS)
al = lambda x: (x + 255) // 256 * 256
off = 0
def take(n):
    global off
    o = off; off += al(n); return o
o_xbar = take(BV * C * 4); o_xs = take(2 * BV * C * 2); o_q = take(2 * BV * EMB * 2); o_wpl = take(BV * 2 * WPLANE * 2)
o_ct = take(BV * HEADS * TP * 4); o_ya = take(2 * BV * HEADS * YA * 2); o_z = take(2 * BV * EMB * 2); o_o = take(BV * EMB * 4)
assert off == need, (off, need)

def setup(kernel):
    """weights folded for `kernel` ('mma' | 'umma'), its channel orders and w_eff plane pitch"""
    global w, score_ch, sum_ch, WPITCH, WPLANE
    os.environ["PT_POOL_KERNEL"] = kernel
    w = m._weights(img.device)
    score_ch, sum_ch = ops.img_pool_channel_orders(img.device, w["img"]["variant"])
    WPITCH = 512 if kernel == "umma" else 528
    WPLANE = HEADS * WPITCH

def run(single=False):
    ws = torch.zeros(need + BV * DBG * 4, dtype=torch.uint8, device="cuda")
    out, ws = ops.img_attnpool(img, w["img"], HEADS, params=w["img_struct"], stages=1, ws=ws)
    out, ws = ops.img_attnpool(img, w["img"], HEADS, params=w["img_struct"], stages=2, out=out, ws=ws)
    torch.cuda.synchronize()
    return out, ws

def f32(ws, o, n): return ws[o:o + 4 * n].view(torch.float32)
def bf(ws, o, n): return ws[o:o + 2 * n].view(torch.bfloat16)

KERNELS = [k for k in ("mma", "umma") if k in sys.argv] or ["mma", "umma"]
MODES = [(k, False) for k in KERNELS]
finals = {}
for kernel, single in MODES:
    setup(kernel)
    out, ws = run(single)
    finals["single" if single else kernel] = out.clone()
    X = img.reshape(BV, C, HW).double()
    wpl = bf(ws, o_wpl, BV * 2 * WPLANE).double().reshape(BV, 2, HEADS, WPITCH)[..., :512].sum(1)      # (BV, 8, 512) score order
    weff = torch.zeros(BV, HEADS, C, dtype=torch.float64, device="cuda")
    weff[:, :, score_ch] = wpl
    cterm = f32(ws, o_ct, BV * HEADS * TP).double().reshape(BV, HEADS, TP)
    xbar = f32(ws, o_xbar, BV * C).double().reshape(BV, C)
    sc = torch.cat([torch.einsum("vhc,vc->vh", weff, xbar)[..., None], torch.einsum("vhc,vct->vht", weff, X)], -1)   # (BV,8,226)
    sv = (sc + cterm[..., :226]) / np.sqrt(32.0)
    P = torch.softmax(sv, -1)
    Y = torch.einsum("vht,vct->vhc", P[..., 1:], X) + P[..., :1] * xbar[:, None, :]
    dbg = f32(ws, need, BV * DBG).reshape(BV, DBG)
    sv_k = dbg[:, :HEADS * 256].reshape(BV, HEADS, 256)[..., :226].double()
    ya = bf(ws, o_ya, 2 * BV * HEADS * YA).double().reshape(2, BV, HEADS, YA).sum(0)
    P_k = ya[..., 512:512 + 226]
    Y_k = torch.zeros_like(Y); Y_k[:, :, sum_ch] = ya[..., :512]
    name = "single" if single else kernel
    print(f"[{name}] scores max err {float((sv_k - sv).abs().max()):.3e}  probs max err {float((P_k - P).abs().max()):.3e}  "
          f"sums max rel err {float(((Y_k - Y).abs().max() / Y.abs().max())):.3e}")
    d = (sv_k - sv).abs()
    print(f"[{name}] score err by (tau mod 8):", [f"{float(d[..., 1:][..., r::8].max()):.2e}" for r in range(8)], "token0", f"{float(d[..., 0].max()):.2e}")
    print(f"[{name}] worst view score err {float(d.amax(dim=(1, 2)).max()):.2e} (view {int(d.amax(dim=(1, 2)).argmax())})")
    dy = (Y_k - Y).abs().amax(dim=(0, 1)).reshape(8, 64)
    print(f"[{name}] sums err by slab:", [f"{float(dy[k].max()):.1e}" for k in range(8)])

if len(finals) > 1:
    ks = list(finals)
    print("final image proxies, max |%s - %s| = %.3e" % (ks[0], ks[-1], float((finals[ks[0]] - finals[ks[-1]]).abs().max())))

if len(sys.argv) > 3 and sys.argv[3] not in ("single", "mma", "umma"):                 # timing: BACK stage (pool + value GEMMs + LayerNorm) at bench size
    Bt, Vt = 64, 196
    imgs = [(torch.relu(torch.randn(Bt, Vt, C, 15, 15, device="cuda")) * 1.5).bfloat16() for _ in range(2)]   # 2 x 2.9 GB >> L2
    needt = _lib.load().pt_img_attnpool_ws_bytes(Bt * Vt, C, HW, EMB, HEADS)
    wss = [torch.zeros(needt, dtype=torch.uint8, device="cuda") for _ in range(2)]
    outs_t = [ops.img_attnpool(imgs[k], w["img"], HEADS, params=w["img_struct"], stages=1, ws=wss[k])[0] for k in range(2)]
    for pf in ([int(x) for x in sys.argv[3].split(',')] if sys.argv[3][0].isdigit() else (4, 0, 2, 6, 8, 10, 4)):
        os.environ["PT_POOL_PF"] = str(pf)
        setup("mma")
        for k in range(4):
            ops.img_attnpool(imgs[k & 1], w["img"], HEADS, params=w["img_struct"], stages=2, out=outs_t[k & 1], ws=wss[k & 1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(40):
            ops.img_attnpool(imgs[k & 1], w["img"], HEADS, params=w["img_struct"], stages=2, out=outs_t[k & 1], ws=wss[k & 1])
        e1.record(); torch.cuda.synchronize()
        print(f"BACK stage (pool + value GEMMs + LayerNorm), PT_POOL_PF={pf}: {e0.elapsed_time(e1) / 40:.4f} ms per {Bt} scenes")

"""Tiny driver for ncu captures of the contraction kernel: a few launches of one shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aces4_b200 as sip
api = sip.api
sip.init(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "gemm"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
if mode == "gemm":
    A, B, C = api.DeviceBlock((n, n)).fill(0.5), api.DeviceBlock((n, n)).fill(0.25), api.DeviceBlock((n, n))
    for _ in range(3):
        api.dgemm_tn(n, n, n, A, n, B, n, C, n)
elif mode == "ring":
    v, o = 50, 20
    nb = n
    Ls = [api.DeviceBlock((v, o, v, o)).fill(0.5) for _ in range(8)]
    Rs = [api.DeviceBlock((v, o, v, o)).fill(0.25) for _ in range(8)]
    Ds = [api.DeviceBlock((v, o, v, o)) for _ in range(nb)]
    ptrn, _ = api.get_contraction_ptrn([1, 2, 3, 4], [1, 2, 5, 6], [5, 6, 3, 4])
    bc = api.BatchedContraction(ptrn, [(v, o, v, o)] * nb, [(v, o, v, o)] * nb, [(v, o, v, o)] * nb,
                                [Ls[i % 8].ptr for i in range(nb)], [Rs[(i * 3) % 8].ptr for i in range(nb)], [d.ptr for d in Ds])
    for _ in range(3):
        bc.launch()
elif mode == "small":
    # block-sparse batching regime: nb blocks of s^4 (GEMM (s^2)^3) in one launch
    s_ = n
    nb = 37 * 24
    Ls = [api.DeviceBlock((s_,) * 4).fill(0.5) for _ in range(8)]
    Rs = [api.DeviceBlock((s_,) * 4).fill(0.25) for _ in range(8)]
    Ds = [api.DeviceBlock((s_,) * 4) for _ in range(nb)]
    ptrn, _ = api.get_contraction_ptrn([1, 2, 3, 4], [1, 2, 5, 6], [5, 6, 3, 4])
    bc = api.BatchedContraction(ptrn, [(s_,) * 4] * nb, [(s_,) * 4] * nb, [(s_,) * 4] * nb,
                                [Ls[i % 8].ptr for i in range(nb)], [Rs[(i * 3) % 8].ptr for i in range(nb)], [d.ptr for d in Ds])
    for _ in range(3):
        bc.launch()
elif mode in ("skinny", "outer"):
    # skinny: D[a,i,b,j] = L[a,i,c,j]*R[c,b] (68 of the 170 SIAL patterns: one segment-sized free + contracted index)
    # outer : D[a,b] = L[a,i,c,j]*R[b,i,c,j] (rank-2 result of two rank-4 blocks: K = 20000, M = N = 50)
    v, o, nb = 50, 20, n
    if mode == "skinny":
        dl, ll, rl = [1, 2, 3, 4], [1, 2, 5, 4], [5, 3]
        lsh, rsh, dsh = (v, o, v, o), (v, v), (v, o, v, o)
    else:
        dl, ll, rl = [1, 2], [1, 3, 4, 5], [2, 3, 4, 5]
        lsh, rsh, dsh = (v, o, v, o), (v, o, v, o), (v, v)
    Ls = [api.DeviceBlock(lsh).fill(0.5) for _ in range(min(nb, 128))]
    Rs = [api.DeviceBlock(rsh).fill(0.25) for _ in range(min(nb, 128))]
    Ds = [api.DeviceBlock(dsh) for _ in range(min(nb, 128))]
    ptrn, _ = api.get_contraction_ptrn(dl, ll, rl)
    bc = api.BatchedContraction(ptrn, [lsh] * nb, [rsh] * nb, [dsh] * nb, [Ls[i % len(Ls)].ptr for i in range(nb)],
                                [Rs[i % len(Rs)].ptr for i in range(nb)], [Ds[i % len(Ds)].ptr for i in range(nb)])
    for _ in range(3):
        bc.launch()
elif mode == "permute16":
    shape = (16, 16, 16, 16)
    ins = [api.DeviceBlock(shape).fill(1.0) for _ in range(1024)]
    outs = [api.DeviceBlock(shape) for _ in range(1024)]
    for transp in ([1, 4, 3, 2, 1], [1, 2, 1, 4, 3]):
        for _ in range(2):
            api.permute_batched(ins, transp, outs)
elif mode == "permute":
    shape = (64, 64, 64, 64) if n >= 64 else (50, 20, 50, 20)
    a, b = api.DeviceBlock(shape).fill(1.0), api.DeviceBlock(shape)
    for transp in ([1, 3, 4, 1, 2], [1, 2, 1, 4, 3], [1, 4, 3, 2, 1]):
        for _ in range(2):
            api.permute(a, transp, out=b)
sip.sync()

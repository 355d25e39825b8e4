"""tcgen05 building block: the TMEM-A / smem-B UMMA with the operand layouts of the fused kernel."""
import pytest
import torch

from pesto_b200 import _lib

pytestmark = pytest.mark.gpu


def probe(A, B, split, lbo=-1, sbo=-1, idesc=0):
    lib = _lib.load()
    K, N = A.shape[1], B.shape[0]
    D = torch.full((128, N), float("nan"), device="cuda")
    _lib.check(lib.pesto_debug_umma_probe(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, int(split), lbo, sbo, idesc, None), "probe")
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("K,N", [(32, 128), (32, 32), (64, 64), (32, 16), (128, 128), (96, 48)])
def test_umma_matches_matmul(K, N):
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = (torch.randn(128, K, generator=g) * 3).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    ref64 = (A.double() @ B.double().T)
    # single pass: operands rounded to the library's 16-bit plane format (fp16 by default, bf16 with -DPESTO_SPLIT_BF16),
    # fp32 accumulation
    D1 = probe(A, B, 0)
    refs = [(A.to(t).double() @ B.to(t).double().T) for t in (torch.float16, torch.bfloat16)]
    assert min((D1.double() - r).abs().max().item() / r.abs().max().item() for r in refs) < 1e-3
    # 3-term split: close to the fp32 product (fp16 planes: ~2^-21, bf16 planes: ~2^-16 relative)
    D3 = probe(A, B, 1)
    assert (D3.double() - ref64).abs().max().item() < 2e-5 * ref64.abs().max().item()


def test_edge_timeline_debug_entry():
    """pesto_debug_edge_timeline: CTA 0 of the tensor-core edge kernel records 19 clock stamps per tile; the first 17
    are phase boundaries in program order, so they must be non-decreasing, and switching it off must stop the writes."""
    import json
    import os
    import numpy as np
    from conftest import GOLDEN
    from pesto_b200.model import Model
    from pesto_b200.data_encoding import extract_topology
    from pesto_b200.synth import synth_structure, one_hot_features
    lib = _lib.load()
    with open(os.path.join(GOLDEN, "config_i_v4_0.json")) as fh:
        model = Model(json.load(fh), mode="bf16x3")
    model.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, "weights_i_v4_0.npz")).items()})
    model = model.eval().cuda()
    X, el, rid = synth_structure(4096, 7)
    Xd = X.cuda()
    ids1 = extract_topology(Xd, 64)[0] + 1
    q0, ridd = one_hot_features(el).cuda(), rid.int().cuda()
    tiles = 4
    buf = torch.zeros((tiles, 2, 2, 19), dtype=torch.int64, device="cuda")
    _lib.check(lib.pesto_debug_edge_timeline(buf.data_ptr(), tiles), "timeline on")
    with torch.no_grad():
        z = model(Xd, ids1, q0, ridd, n_res=int(rid.max()) + 1)
    torch.cuda.synchronize()
    _lib.check(lib.pesto_debug_edge_timeline(None, 0), "timeline off")
    t = buf.cpu().numpy()
    assert (t[..., :17] > 0).all() and torch.isfinite(z).all()
    assert (np.diff(t[..., :17], axis=-1) >= 0).all()            # program order inside a tile
    assert (np.diff(t[..., 0], axis=0) > 0).all()                # tiles of one pipeline follow each other
    buf.zero_()
    with torch.no_grad():
        model(Xd, ids1, q0, ridd, n_res=int(rid.max()) + 1)
    torch.cuda.synchronize()
    assert int(buf.abs().sum()) == 0


@pytest.mark.parametrize("issue_lanes", [1, 4, 32, 128])
def test_rmma_probe_gather4_and_mn_major_operands(issue_lanes):
    """TMA tile::gather4 (64-byte swizzle) of fp16 hi|lo neighbour records + tcgen05.mma with both operands MN-major
    in shared memory (A = the records as they landed, B = per-edge weights written by threads): the building blocks
    of the edge kernel's tensor-core reduction, against a float64 matmul."""
    import os
    import numpy as np
    lib = _lib.load()
    g = torch.Generator().manual_seed(11)
    n_rows = 1000
    p = torch.randn(n_rows, 96, generator=g) * 3
    p[0] = 0
    hi = p.half()
    lo = (p - hi.float()).half()
    p16 = torch.cat([hi, lo], 1).contiguous().cuda()
    ids = torch.randint(0, n_rows, (128,), generator=g, dtype=torch.int32)
    ids[5] = 0
    ids[77] = n_rows - 1
    W = torch.randn(16, 128, generator=g)
    D = torch.full((128, 16), float("nan"), device="cuda")
    Prec = torch.full((128, 96), float("nan"), device="cuda")
    raw = torch.zeros(48 * 1024 // 4, dtype=torch.int32, device="cuda")
    status = torch.zeros(3, dtype=torch.int32, device="cuda")
    ids_d, W_d = ids.cuda(), W.cuda()                                  # (named: the pointers must outlive the call)
    _lib.check(lib.pesto_debug_rmma_probe(p16.data_ptr(), n_rows, ids_d.data_ptr(), W_d.data_ptr(), D.data_ptr(),
                                          Prec.data_ptr(), raw.data_ptr(), -1, -1, -1, -1, 0, issue_lanes, status.data_ptr(), None), "rmma probe")
    torch.cuda.synchronize()
    print(f"gather of 128 x 384 B rows, {issue_lanes} issuing lane(s): issue {int(status[1])} cycles, landed after {int(status[2])} cycles")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        np.savez(os.path.join(out, "rmma_probe.npz"), D=D.cpu().numpy(), Prec=Prec.cpu().numpy(), raw=raw.cpu().numpy(),
                 p16=p16.cpu().numpy().view(np.uint16), ids=ids.numpy(), W=W.numpy(), status=status.cpu().numpy())
    assert int(status[0]) == 0, f"probe wait timed out at stage {int(status[0])}"
    pg = (hi.double() + lo.double())[ids.long()]                       # [128 edges][96]
    assert (Prec.cpu().double() - pg).abs().max().item() < 1e-6        # thread-side read of the swizzled tile
    ref = pg.T @ W.double().T                                          # [96][16]
    err = (D.cpu().double()[:96] - ref).abs().max().item()
    assert err < 2e-5 * ref.abs().max().item(), err

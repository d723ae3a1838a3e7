"""One small invocation of the hot path on cuda:0, checked against the CPU oracle (used by
__graft_entry__.smoke()): an IPA layer-call and one teacher-forced SE(3)/categorical reverse step."""
import json
import os
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run():
    from abx_b200 import lib
    from abx_b200.diffuser.full_diffuser import FullDiffuser
    from abx_b200.model.folding import InvariantPointAttention
    from abx_b200.utils.weights import np_randn, seeded_state_dict
    from oracle import diffusers as OD
    from oracle import model as OM
    from oracle import quat as OQ

    lib.reset_launch_count()
    dev = torch.device('cuda:0')
    shapes = {k: tuple(v) for k, v in json.load(open(os.path.join(ROOT, 'tests', 'golden', 'state_dict_shapes.json'))).items()}
    P = seeded_state_dict(shapes, 0)
    prefix = OM.SN + 'attention_module.'

    # ---- IPA layer-call ------------------------------------------------------------------------------
    B, N = 2, 48
    ipa = InvariantPointAttention(OM.IPA_CONF, 128)
    ipa.load_state_dict({k[len(prefix):]: v for k, v in P.items() if k.startswith(prefix)})
    ipa = ipa.to(dev).eval()
    g = torch.Generator().manual_seed(0)
    x, z = np_randn(1, B, N, 256), np_randn(2, B, N, N, 128)
    q = torch.randn(B, N, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    rots, trans = OQ.quat_to_rot(q), torch.randn(B, N, 3, generator=g) * 2
    mask = torch.ones(B, N)
    mask[1, -7:] = 0
    ref = OM.ipa_forward(P, x, z, mask, rots, trans)
    with torch.no_grad():
        out = ipa(x.to(dev), z.to(dev), mask.to(dev), (rots.to(dev), trans.to(dev))).cpu()
    err = float((out - ref).abs().max() / ref.abs().max())
    assert err < 3e-5, f'IPA mismatch vs oracle: {err}'

    # ---- reverse step ----------------------------------------------------------------------------------
    cfg = json.load(open(os.path.join(ROOT, 'abx_b200', 'config', 'config_model.json')))['diffuser']
    cfg['so3'].update(use_cached_score=True, cache_dir=tempfile.mkdtemp(prefix='abx_smoke_'), num_sigma=200, num_omega=200)
    fd = FullDiffuser(cfg)
    conf = json.loads(json.dumps(OD.DEFAULT_CONF))
    conf['so3'].update(num_sigma=200, num_omega=200)
    od = OD.OracleDiffuser(fd._so3_diffuser._score_norms, conf=conf)
    rig = torch.cat([q, trans * 10], -1).double()
    seq = torch.randint(0, 20, (B, N), generator=g)
    rot_score, logits = torch.randn(B, N, 3, generator=g), torch.randn(B, N, 20, generator=g)
    trans_score = torch.randn(B, N, 3, generator=g).double()
    dmask = (torch.rand(B, N, generator=g) < 0.4).int()
    t = torch.full((B,), 0.5, dtype=torch.float64)
    dt = torch.tensor(0.01)
    z_rot, z_tr = torch.randn(B, N, 3, generator=g), torch.randn(B, N, 3, generator=g)
    jumps = torch.poisson(od.reverse_rates(seq, logits, t)[0] * dt, generator=g)
    r_ref, s_ref = od.reverse(rig, seq, rot_score, trans_score, logits, t, dt, dmask, z_rot, z_tr, jumps)
    c = lambda v: v.to(dev)  # noqa: E731
    r_out, s_out = fd.reverse(c(rig), c(seq), c(rot_score), c(trans_score), c(logits), c(t), dt, diffuse_mask=c(dmask),
                              noise=(c(z_rot), c(z_tr), c(jumps)))
    assert torch.equal(s_out.cpu(), s_ref.long()), 'residue types differ from the oracle'
    err2 = float((r_out.cpu() - r_ref).abs().max())
    assert err2 < 1e-9, f'reverse step mismatch vs oracle: {err2}'
    print(f'smoke ok: ipa rel err {err:.2e}, reverse abs err {err2:.2e}, abx kernels launched {lib.launch_count()}')

"""The functional oracle of the score network (oracle/model.py) against outputs of the reference's
own modules (tests/golden/{ipa,ipascore,model,sampler}.npz).  CPU only."""
import torch

from abx_b200.utils.weights import np_randn
from oracle import model as M
from tests.util import batch_from_golden, golden, maxabs, oracle_diffuser, seeded_params


def test_ipa_matches_reference():
    g = golden('ipa')
    P = seeded_params()
    x, z = np_randn(211, 2, 37, 256), np_randn(212, 2, 37, 37, 128)
    out = M.ipa_forward(P, x, z, g['mask'], g['rots'], g['trans'])
    assert maxabs(out, g['out']) < 2e-5 * float(g['out'].abs().max())


def test_ipascore_matches_reference():
    g = golden('ipascore')
    P = seeded_params()
    batch = batch_from_golden(g)
    B, N = batch['seq'].shape
    out = M.ipascore_forward(P, oracle_diffuser(), np_randn(221, B, N, 544), np_randn(222, B, N, N, 192), batch)
    assert maxabs(out['structure_module'], g['structure_module']) < 1e-4
    assert maxabs(out['rigids'], g['rigids']) < 1e-4
    assert maxabs(out['angles_sin_cos'], g['angles_sin_cos']) < 1e-4
    assert maxabs(torch.stack([t for _, t in out['traj']]), g['traj_trans']) < 1e-4
    assert maxabs(out['trans_score'], g['trans_score']) < 1e-4
    assert maxabs(out['rot_score'], g['rot_score']) < 1e-3 * max(1.0, float(g['rot_score'].abs().max()))


def test_score_network_matches_reference():
    g = golden('model')
    P = seeded_params()
    batch = batch_from_golden(g)
    B, N = batch['seq'].shape
    b1 = dict(batch)
    b1.update(prev_pos=torch.zeros(B, N, N, dtype=torch.int64), prev_seq=torch.zeros(B, N, 544),
              prev_pair=torch.zeros(B, N, N, 192))
    s, p = M.embed_inputs(P, b1)
    s, p = M.seqformer_block(P, s, p, b1['mask'])
    assert maxabs(s, g['trunk_seq']) < 1e-4 * max(1.0, float(g['trunk_seq'].abs().max()))
    assert maxabs(p[:, :8, :8], g['trunk_pair']) < 1e-4 * max(1.0, float(g['trunk_pair'].abs().max()))

    out = M.score_network(P, oracle_diffuser(), batch)
    assert torch.equal(batch['seq_t'], g['seq_t_after'])           # recycling overwrote seq_t (abx.py:97-98)
    assert torch.equal(out['seq_0'], g['seq_0'])
    assert maxabs(out['rigids'], g['rigids']) < 1e-3
    assert maxabs(out['atom14'], g['atom14']) < 1e-3
    assert maxabs(out['atom37'], g['atom37']) < 1e-3
    assert maxabs(out['trans_score'], g['trans_score']) < 1e-3
    assert maxabs(out['logits'], g['logits']) < 1e-3
    assert maxabs(out['pLDDT'], g['pLDDT']) < 1e-2
    assert torch.equal(M.get_prev(batch, out)['prev_pos'], g['prev_pos'])


def test_sampler_steps_match_reference():
    """Warm-up self-conditioning call + two teacher-forced loop iterations (t = 1.0, 0.99)."""
    from oracle import sampler as S
    g = golden('sampler')
    P = seeded_params()
    od = oracle_diffuser()
    batch = batch_from_golden(g)
    grid = S.reverse_grid()
    dt = torch.tensor(1 / 100)
    batch = S.self_condition(P, od, batch, grid[0])
    for k in range(2):
        out = S.sample_step(P, od, batch, grid[k], dt, (g[f's{k}_z_rot'], g[f's{k}_z_trans'], g[f's{k}_jumps']))
        assert maxabs(out['trans_score'], g[f's{k}_trans_score']) < 1e-3
        assert maxabs(out['logits'], g[f's{k}_logits']) < 1e-3
        assert maxabs(out['atom14'], g[f's{k}_atom14']) < 1e-3
        assert torch.equal(batch['seq_t'].long(), g[f's{k}_seq'].long())          # bit-exact residue types
        assert batch['rigids_t'].dtype == torch.float64
        assert maxabs(batch['rigids_t'], g[f's{k}_rigids']) < 1e-4               # frames within 1e-4

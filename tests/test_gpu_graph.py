"""GraphedTrainStep (dl_vqa_b200/graph.py): the loop body of reference train.py:69-81 captured as one CUDA graph, with
the per-step host arithmetic (dropout seed, LR schedule of train.py:31-35, Adam bias corrections) in a device-resident
VqaStepState.  The replayed step must follow the same trajectory as the kernel-by-kernel step."""
import pytest
import torch

from oracle import vqa_oracle as O

pytestmark = pytest.mark.gpu


def _setup(dtype, dropout, seed=4):
    import dl_vqa_b200 as D
    cfg = O.cfg_with(O.DEFAULT_CFG, image_size=64)
    cfg = O.zero_dropout(cfg) if dropout == 0 else cfg
    V = 300
    sd = O.random_params(cfg, V, seed=seed, scale=1.5)
    v, q, q_len, a_idx, a_val, a_len = O.synthetic_batch(6, cfg, V, seed=8, T=9)
    m = D.VqaNet(cfg, V, compute_dtype=dtype)
    m.load_state_dict(sd)
    m.cuda().train(True)
    batch = tuple(t.cuda() for t in (v, q, a_idx, a_val, a_len)) + (None, q_len.cuda())
    return D, cfg, m, batch


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_graph_replay_follows_the_eager_trajectory(dtype):
    """Dropout 0: six steps through GraphedTrainStep (1 eager + 1 capture/replay + 4 replays) against six steps of the plain
    loop (run_batch, zero_grad, update_learning_rate, backward, FusedAdam.step) from identical weights."""
    D, cfg, m_ref, batch = _setup(dtype, 0)
    lr0 = 1e-3
    opt_ref = D.FusedAdam(m_ref.parameters(), lr=lr0)
    want = []
    for it in range(6):
        loss, score = D.run_batch(m_ref, None, batch, cfg["max_answers"])
        opt_ref.zero_grad(set_to_none=True)
        D.update_learning_rate(opt_ref, it, lr0)
        loss.backward()
        opt_ref.step()
        want.append(float(loss.detach()))

    D, cfg, m, batch = _setup(dtype, 0)
    opt = D.FusedAdam(m.parameters(), lr=lr0)
    if dtype == "bfloat16":
        m.use_weight_shadows(opt)
    step = D.GraphedTrainStep(m, opt, cfg["max_answers"], lr=lr0)
    got = []
    for it in range(6):
        loss, score = step(batch)
        got.append(float(loss))                  # read before the next replay overwrites the static output
    assert step.replays == 5 and step.launches_per_replay > 20
    # bf16: summation-order noise (atomics) flips bf16 roundings and the two runs drift apart step by step, as two eager runs
    # do; the first steps pin the mechanics (same lr, same bias corrections, same weights read), the later ones the trend
    for i, (a, b) in enumerate(zip(got, want)):
        tol = 1e-4 if dtype == "float32" else (5e-3 if i < 3 else 5e-2)
        assert abs(a - b) <= tol * abs(b), (i, got, want)
    assert got[-1] < got[0]
    st = step.read_state()
    assert st["iteration"] == 6 and st["adam_step"] == 6
    assert abs(st["lr"] - O.learning_rate(lr0, 5)) <= 1e-7 * lr0            # lr of the LAST executed step (iteration 5)
    bc1 = 1 - 0.9 ** 6
    assert abs(st["lr_over_bc1"] - O.learning_rate(lr0, 5) / bc1) <= 1e-6 * lr0
    # host mirrors stay meaningful
    assert all(int(opt.state[p]["step"]) == 6 for p in m.parameters())
    for pa, pb in zip(m.parameters(), m_ref.parameters()):
        # Adam normalises every gradient entry by its own magnitude: entries that are pure summation-order noise (atomics)
        # move by +-lr per step in either run, so parameters agree to a few lr, not to 1e-4; the loss trajectory above is
        # the tight check
        assert O.rel_err(pa, pb) < 6e-2


def test_graph_replay_draws_a_fresh_dropout_mask_every_step_and_lr_decays_on_device():
    """Dropout 0.3, lr 0 (weights frozen): the loss of a replayed graph on the SAME inputs must change from replay to replay
    (the seed is read from device memory at run time) while eval-mode logits stay put."""
    D, cfg, m, batch = _setup("bfloat16", 0.3)
    opt = D.FusedAdam(m.parameters(), lr=0.0)
    step = D.GraphedTrainStep(m, opt, cfg["max_answers"], lr=0.0, half_life=2.0)
    losses, seeds = [], []
    for _ in range(6):
        loss, _ = step(batch)
        losses.append(float(loss))
        seeds.append(step.read_state()["seed"])
    assert step.replays == 5
    assert len(set(seeds)) == 6
    assert len({round(x, 6) for x in losses[1:]}) >= 4, losses          # replays 1..5 differ from one another
    assert max(losses) - min(losses) < 0.5 * abs(losses[0])              # ...but it is the same function in expectation


def test_two_input_buffer_sets_give_two_graphs_sharing_one_pool():
    D, cfg, m, batch = _setup("bfloat16", 0)
    opt = D.FusedAdam(m.parameters(), lr=1e-3)
    step = D.GraphedTrainStep(m, opt, cfg["max_answers"], lr=1e-3)
    other = tuple(t.clone() if torch.is_tensor(t) else t for t in batch)
    out = []
    for i in range(8):
        loss, _ = step(batch if i % 2 == 0 else other)
        out.append(float(loss))
    assert len(step._graphs) == 2 and step.replays == 6      # 2 eager first-sightings, then capture + replay
    assert out[-1] < out[0]
    assert all(x == x for x in out)                                       # no NaN


def test_backward_after_parameter_update_raises():
    """the backward re-reads live weights; an optimizer step between forward and backward must raise, as torch's
    own ops do ('modified by an inplace operation')."""
    D, cfg, m, batch = _setup("bfloat16", 0)
    opt = D.FusedAdam(m.parameters(), lr=1e-3)
    loss0, _ = D.run_batch(m, None, batch, cfg["max_answers"])
    loss0.backward()
    loss1, _ = D.run_batch(m, None, batch, cfg["max_answers"])
    opt.step()                                    # parameters change between loss1's forward and its backward
    with pytest.raises(RuntimeError, match="modified in place"):
        loss1.backward()


def test_score_is_non_differentiable_and_eval_score_lands_on_the_host():
    """the score must not carry a grad_fn (train.py:87 accumulates it for a whole epoch); under no_grad
    (train.py:144 evaluate) it is a CPU tensor like the reference's batch_accuracy result, so `score += batch_score` with
    score = torch.tensor(0.0) (train.py:155,165) works unchanged."""
    D, cfg, m, batch = _setup("float32", 0)
    loss, score = D.run_batch(m, None, batch, cfg["max_answers"])
    assert loss.requires_grad and not score.requires_grad and score.grad_fn is None and score.is_cuda
    m.eval()
    with torch.no_grad():
        loss, score = D.run_batch(m, None, batch, cfg["max_answers"])
    assert not score.is_cuda
    acc = torch.tensor(0.0)
    acc += score
    total = 0
    total += loss
    assert float(acc) == float(score)


def test_out_of_range_answer_ids_and_lengths_do_not_touch_foreign_memory():
    """ids outside 1..N are ignored by the loss kernel (the reference raises), q_len > T is clamped."""
    import dl_vqa_b200 as D
    B, N, A = 4, 50, 3
    big = torch.zeros(B + 2, N, device="cuda")
    logits = big[1:B + 1]
    torch.manual_seed(0)
    logits.copy_(torch.randn(B, N))
    a_idx = torch.tensor([[1, 51, 0], [-3, 2, 0], [10 ** 9, 0, 0], [50, 0, 0]])
    a_val = torch.tensor([[2, 5, 0], [4, 3, 0], [7, 0, 0], [1, 0, 0]])
    lg = logits.clone().requires_grad_(True)
    loss, _ = D.soft_target_loss_and_score(lg, a_idx, a_val)
    loss.backward()
    clean_idx = torch.tensor([[1, 0, 0], [0, 2, 0], [0, 0, 0], [50, 0, 0]])
    clean_val = torch.tensor([[2, 0, 0], [0, 3, 0], [0, 0, 0], [1, 0, 0]])
    want = O.soft_target_loss_dense(logits.cpu(), clean_idx, clean_val)
    assert abs(float(loss) - float(want)) < 1e-5
    assert float(big[0].abs().max()) == 0 and float(big[-1].abs().max()) == 0
    D_, cfg, m, batch = _setup("float32", 0)
    v, q, ai, av, al, _, ql = batch
    with torch.no_grad():
        a = m(v, q, ql)
        b = m(v, q, torch.where(ql == q.shape[1], ql + 5, ql))          # full-length rows claim to be longer than T
    assert torch.equal(a, b)

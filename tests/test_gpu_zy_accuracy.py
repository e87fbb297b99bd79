"""GPU: top-10 segment-retrieval accuracy parity on the learnable synthetic task (north-star: within +-0.5 pt of the
reference on the same synthetic eval).  Both implementations start from the same state_dict and see the same batches
and spatial-dropout centres for 64 Adam steps; the reference side is the CPU oracle (oracle/bm_oracle.py)."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_top10_accuracy_parity_on_synthetic_task():
    import brainmagick_b200 as bb
    from brainmagick_b200 import functional as BF, synthetic
    from oracle import accuracy_task as at, bm_oracle

    cfg = bm_oracle.Config(in_channels=32, out_channels=24, n_subjects=4, hidden=160, merger_channels=40,
                           initial_linear=48, merger_pos_dim=128)
    task = at.make_task(cfg, n_train=1024, n_eval=2048, T=90, noise=3.0)
    sched = at.batches(1024, 64, 4)
    p0 = bm_oracle.init_state_dict(cfg, seed=3)

    # ---- reference side: CPU oracle ----
    p_ref, loss_ref = at.train_oracle(cfg, p0, task, sched, lr=1e-3)
    acc_ref, _ = at.eval_oracle(cfg, p_ref, task, k=10)
    acc1_ref, _ = at.eval_oracle(cfg, p_ref, task, k=1)

    # ---- CUDA drop-in ----
    model = bb.SimpleConv(
        in_channels=dict(meg=cfg.in_channels), out_channels=cfg.out_channels, hidden=dict(meg=cfg.hidden), depth=cfg.depth,
        dilation_period=5, kernel_size=3, skip=True, subject_layers=True, subject_dim=0, complex_out=True, glu=2,
        glu_context=1, merger=True, initial_linear=cfg.initial_linear, merger_channels=cfg.merger_channels, gelu=True,
        batch_norm=True, merger_pos_dim=cfg.merger_pos_dim, n_subjects=cfg.n_subjects)
    model.load_state_dict(p0)
    model = model.cuda().train()
    clip = bb.ClipLoss().cuda().train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999))
    recs = [synthetic.SyntheticRecording(s, task["positions"][s]) for s in range(cfg.n_subjects)]
    d = {k: v.cuda() for k, v in task["train"].items()}
    losses = []
    for idx, ban in sched:
        idx_d = idx.cuda()
        meg, feats, subj = d["meg"][idx_d], d["feats"][idx_d], d["subj"][idx_d]
        batch = synthetic.SyntheticBatch(meg, subj, [recs[int(s)] for s in task["train"]["subj"][idx]])
        model.merger.ban_centre_override = ban
        opt.zero_grad(set_to_none=True)
        est = model(dict(meg=meg), batch)
        loss = clip(est, feats, torch.ones(len(idx), 1, meg.shape[-1], dtype=torch.bool, device="cuda"))
        loss.backward()
        opt.step()
        losses.append(loss.item())
    BF.check_tc_status()

    model.eval()
    clip.eval()
    e = {k: v.cuda() for k, v in task["eval"].items()}
    ests = []
    with torch.no_grad():
        for i in range(0, len(e["meg"]), 256):
            sl = slice(i, i + 256)
            batch = synthetic.SyntheticBatch(e["meg"][sl], e["subj"][sl], [recs[int(s)] for s in task["eval"]["subj"][sl]])
            ests.append(model(dict(meg=e["meg"][sl]), batch))
    est = torch.cat(ests)
    hits10 = hits1 = 0
    for i in range(0, len(est), 256):
        probs = clip.get_probabilities(est[i:i + 256], e["feats"])
        top = probs.topk(10, dim=1).indices
        truth = torch.arange(i, min(i + 256, len(est)), device="cuda")[:, None]
        hits10 += (top == truth).any(dim=1).sum().item()
        hits1 += (top[:, :1] == truth).sum().item()
    acc, acc1 = hits10 / len(est), hits1 / len(est)
    result = dict(task="synthetic latent retrieval, 2048 held-out segments, 64 Adam steps (lr 1e-3, B=64)",
                  top10_reference_cpu_oracle=acc_ref, top10_cuda=acc, top1_reference_cpu_oracle=acc1_ref, top1_cuda=acc1,
                  final_loss_reference=loss_ref[-1], final_loss_cuda=losses[-1], first_loss_reference=loss_ref[0],
                  first_loss_cuda=losses[0])
    print("\n[accuracy parity]", json.dumps(result))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "accuracy_parity.json"), "w") as f:
            json.dump(result, f, indent=1)
    assert abs(losses[0] - loss_ref[0]) < 1e-4 * max(1.0, abs(loss_ref[0]))
    assert acc_ref > 0.3, "the task must be learnt for the comparison to mean anything"
    assert abs(acc - acc_ref) <= 0.005, (acc, acc_ref)          # +-0.5 pt


def test_top10_accuracy_parity_at_baseline_widths():
    """The same comparison at the BASELINE widths (208 sensors, hidden 320, F = 1024, T = 360, 27 subjects):
    4096 training segments (sensor noise 3), 256 Adam steps of B = 32, 1024 held-out segments with sensor noise 8 (a harder
    evaluation, so that the top-k accuracy of a trained model stays away from 100 %).

    Training at this size is CHAOTIC in the escape time from the initial plateau (loss = ln 32): the three implementations
    (CUDA drop-in, the oracle's PyTorch ops on the GPU, the CPU oracle) agree to 4 digits for ~5 steps and then leave the
    plateau anywhere between step 100 and step 250 (profiles/diag_accuracy.py) -- the reference does not reproduce
    itself across devices either.  Two comparisons therefore:
      (a) SAME WEIGHTS: the model trained here through the CUDA path is evaluated twice, by the CUDA path (through
          brainmagick_b200.retrieval.retrieval_accuracy, the batched evaluation) and by the oracle's PyTorch ops (cuDNN fp32,
          TF32 off): top-10 and top-1 within +-0.5 pt (north star), estimates within 1e-4;
      (b) SAME TRAINING RECIPE: against the accuracy the CPU oracle reached when it trained from the same state on the same
          batches in the build container (tests/golden/accuracy_full_width.json, oracle/make_accuracy_golden.py, ~20 min):
          both must have learnt, the first steps agree, and the top-10 accuracies are reported side by side (the run-to-run
          spread of either implementation is several points: see the asserts)."""
    import brainmagick_b200 as bb
    from brainmagick_b200 import functional as BF, retrieval, synthetic
    from conftest import rel_err
    from oracle import bm_oracle, make_accuracy_golden as mg

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "accuracy_full_width.json")
    with open(path) as f:
        gold = json.load(f)
    spec = gold["spec"]
    cfg, task, sched, p0 = mg.build(spec)
    assert len(sched) == gold["steps"]
    model = bb.SimpleConv(
        in_channels=dict(meg=cfg.in_channels), out_channels=cfg.out_channels, hidden=dict(meg=cfg.hidden), depth=cfg.depth,
        dilation_period=5, kernel_size=3, skip=True, subject_layers=True, subject_dim=0, complex_out=True, glu=2,
        glu_context=1, merger=True, initial_linear=cfg.initial_linear, merger_channels=cfg.merger_channels, gelu=True,
        batch_norm=True, merger_pos_dim=cfg.merger_pos_dim, n_subjects=cfg.n_subjects)
    model.load_state_dict(p0)
    model = model.cuda().train()
    clip = bb.ClipLoss().cuda().train()
    opt = torch.optim.Adam(model.parameters(), lr=spec["lr"], betas=(0.9, 0.999))
    recs = [synthetic.SyntheticRecording(s, task["positions"][s]) for s in range(cfg.n_subjects)]
    d = task["train"]
    mask = torch.ones(spec["batch"], 1, spec["T"], dtype=torch.bool, device="cuda")
    losses = []
    for idx, ban in sched:
        meg, feats, subj = d["meg"][idx].cuda(), d["feats"][idx].cuda(), d["subj"][idx].cuda()
        batch = synthetic.SyntheticBatch(meg, subj, [recs[int(s)] for s in d["subj"][idx]])
        model.merger.ban_centre_override = ban
        opt.zero_grad(set_to_none=True)
        loss = clip(model(dict(meg=meg), batch), feats, mask)
        loss.backward()
        opt.step()
        losses.append(loss.detach())
    losses = [float(v) for v in torch.stack(losses).cpu()]
    BF.check_tc_status()
    model.eval()
    clip.eval()
    e = task["eval"]
    n_eval = len(e["meg"])
    # ---- (a) the trained weights, evaluated by the CUDA path and by the oracle's PyTorch ops on the same GPU ----
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        params = {k: v.detach().clone() for k, v in model.state_dict().items()}
        pos_g = task["positions"].cuda()
        ests, ests_ref = [], []
        with torch.no_grad():
            for i in range(0, n_eval, 256):
                sl = slice(i, i + 256)
                meg, subj = e["meg"][sl].cuda(), e["subj"][sl].cuda()
                batch = synthetic.SyntheticBatch(meg, subj, [recs[int(s)] for s in e["subj"][sl]])
                ests.append(model(dict(meg=meg), batch))
                ests_ref.append(bm_oracle.simpleconv_forward(params, cfg, meg, pos_g, subj, subj, False))
        est, est_ref = torch.cat(ests), torch.cat(ests_ref)
        feats_g = e["feats"].cuda()
        ar = torch.arange(n_eval, device="cuda")
        acc_ref = {k: bm_oracle.topk_accuracy(est_ref, feats_g, ar, k=k, on="scores") for k in (1, 10)}
        acc_ref_probs = {k: bm_oracle.topk_accuracy(est_ref, feats_g, ar, k=k, on="probs") for k in (1, 10)}
        acc_cuda_est_oracle_metric = {k: bm_oracle.topk_accuracy(est, feats_g, ar, k=k, on="scores") for k in (1, 10)}
        tied = bm_oracle.degenerate_rows(est_ref, feats_g, k=10)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    labels = torch.arange(n_eval)
    acc = retrieval.retrieval_accuracy(clip, est, e["feats"], labels, labels, topk=(1, 10), batch_size=256)
    e_est = rel_err(est.cpu(), est_ref.cpu())
    result = dict(task=gold["what"], spec=spec, steps=len(sched),
                  same_weights=dict(top10_cuda=acc[10], top10_oracle_ops=acc_ref[10], top1_cuda=acc[1],
                                    top1_oracle_ops=acc_ref[1], estimate_rel_err=e_est,
                                    top10_oracle_ops_via_fp32_softmax=acc_ref_probs[10],
                                    top10_cuda_estimates_oracle_metric=acc_cuda_est_oracle_metric[10],
                                    rows_decided_by_softmax_ties=tied,
                                    note="ranking on scores; where the fp32 softmax underflows to tied zeros the reference's "
                                         "probs.topk is decided by tie-breaking (rows counted above)"),
                  same_recipe=dict(top10_cuda_trained=acc[10], top10_cpu_oracle_trained=gold["top10"], top1_cuda_trained=acc[1],
                                   top1_cpu_oracle_trained=gold["top1"], first_loss_cuda=losses[0],
                                   first_loss_cpu_oracle=gold["losses"][0], final_loss_cuda=losses[-1],
                                   final_loss_cpu_oracle=gold["losses"][-1],
                                   left_plateau_at_step=dict(cuda=next((i for i, v in enumerate(losses) if v < 2.5), None),
                                                             cpu_oracle=next((i for i, v in enumerate(gold["losses"]) if v < 2.5), None))))
    print("\n[accuracy parity, BASELINE widths]", json.dumps(result))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "accuracy_parity_full_width.json"), "w") as f:
            json.dump(result, f, indent=1)
    assert abs(losses[0] - gold["losses"][0]) < 1e-4 * max(1.0, abs(gold["losses"][0]))
    assert e_est < 1e-4
    # (a) same weights.  The two implementations of the METRIC agree on identical estimates (a near-tie may flip a segment or
    # two), and the two implementations of the MODEL give the same accuracy: +-0.5 pt is the north-star bar.
    assert abs(acc[10] - acc_cuda_est_oracle_metric[10]) <= 2.0 / n_eval, (acc, acc_cuda_est_oracle_metric)
    assert abs(acc[10] - acc_ref[10]) <= 0.005 and abs(acc[1] - acc_ref[1]) <= 0.005, (acc, acc_ref)
    if tied == 0:
        assert abs(acc[10] - acc_ref_probs[10]) <= 0.005, (acc, acc_ref_probs)
    # (b) same recipe, independent runs.  The trajectories agree while rounding differences are still small, then the
    # plateau escape amplifies them and the model goes on to over-fit the 4096 training segments (final loss ~0), so the
    # held-out accuracy depends on the path taken: FOUR runs of THIS implementation, whose fp64 statistics atomics commit in a
    # different order each time, ended at 43.5 %, 42.1 %, 37.0 % and 27.6 % top-10; the CPU oracle at 48.1 %.  Asserted: the
    # first steps agree and both sides learnt (chance is 1 %); the accuracies are REPORTED side by side, the +-0.5 pt claim is
    # carried by (a).
    for i in range(3):
        assert abs(losses[i] - gold["losses"][i]) < 2e-3 * max(1.0, abs(gold["losses"][i])), (i, losses[i], gold["losses"][i])
    assert gold["top10"] > 0.15 and acc[10] > 0.05, "both sides must have learnt the task (chance: 0.01)"

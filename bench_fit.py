#!/usr/bin/env python
"""Fit-loop benchmark for BASELINE.json configs[2] and configs[3] (not part of the bench.py contract).

    python bench_fit.py                       # config 3: 1 image, 20 k object samples, 300 Adam iterations, 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench_fit.py --batch 32               # config 4: 32 images sharded over N GPUs (4 per GPU at N = 8)

Per rank (images are independent, SURVEY.md section 8e: no data-path collective):
  1. CHORE.filter on the rank's images (pinned host -> device copy inside the timed region),
  2. one 20 k-point query per image of all four heads (the neural point cloud stage's kernel),
  3. `iters` fit iterations, one iteration = SMPL-phase step (LBS -> landmarks -> query 6890 verts -> every
     forward_smpl term -> adjoints -> Adam) + 'object only' step (SO(3) -> rigid `points` samples -> query ->
     object/scale/ocent -> adjoints -> Adam) for all of the rank's images at once: FusedFitSteps replayed from
     CUDA graphs (recon/recon_fit_behave.py:90-163,224-337),
  4. NCCL all-gather of the fitted parameters (pose 156 + betas 10 + trans 3 + R 9 + t 3 + s 1 floats per image).
Timing: CUDA events around 1-4 on every rank, max over ranks; prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def make_fit_problem(net, dev, B, n_obj, seed=100):
    """Synthetic fitting problem with the true shapes: SMPL-H-shaped body model, landmark
    regressors (137 x 6890, ~270 non-zeros per row), Mahalanobis priors, part labels, 2-D keypoints, `n_obj` object
    samples per image.  Returns (fitter, crop_center, build_state) where build_state() makes fresh parameters and a
    FusedFitSteps over them (phase 'kpts': every forward_smpl term)."""
    import scipy.sparse as sp
    import chore_b200
    from chore_b200.fitter import HandPrior, MahalanobisPrior
    from oracle import chore_oracle as O          # input synthesis only
    dev = torch.device(dev)
    layer = chore_b200.SMPLHLayer(O.make_smplh_buffers(0), device=str(dev))
    g = torch.Generator().manual_seed(seed)
    regs = [sp.random(L, 6890, density=270 / 6890, format="csr", random_state=7 + i, dtype="float32") for i, L in enumerate((25, 70, 42))]
    mk_prec = lambda n: torch.tril(0.3 * torch.randn(n, n, generator=g)) + torch.eye(n)
    fit = chore_b200.ReconFitterBehave(device=str(dev), strict=True, priors=(
        MahalanobisPrior(0.1 * torch.randn(63, generator=g), mk_prec(63), device=str(dev)),
        HandPrior(0.1 * torch.randn(90, generator=g), mk_prec(45), mk_prec(45), device=str(dev))))
    cc = torch.tensor([[1008., 995.]], device=dev).repeat(B, 1)
    labels = torch.randint(14, (B, 6890), generator=g).to(dev)
    obj0 = (0.2 * torch.randn(B, n_obj, 3, generator=g)).to(dev)
    pose0 = 0.1 * torch.randn(B, 156, generator=g)
    kpts = torch.cat([512 * torch.rand(B, 25, 2, generator=g), torch.rand(B, 25, 1, generator=g)], -1).to(dev)
    data = {"net": net, "query_dict": {"crop_center": cc}, "part_labels": labels, "objects": obj0,
            "pose_init": pose0[:, 3:72].to(dev), "body_kpts": kpts,
            "smpl_center": torch.tensor([[0.0, 0.1, 2.2]], device=dev).repeat(B, 1)}

    def build_state():
        smpl = chore_b200.SMPLPyTorchWrapperBatch(layer, B, betas=0.3 * torch.randn(B, 10, generator=g), pose=pose0.clone(),
                                                  trans=torch.tensor([[0.0, 0.1, 2.2]]).repeat(B, 1), device=str(dev), regressors=regs)
        split = fit.split_smpl(smpl)
        R = (torch.eye(3).repeat(B, 1, 1) + 0.05 * torch.randn(B, 3, 3, generator=g)).to(dev).requires_grad_(True)
        t = torch.tensor([[0.2, 0.1, 2.3]], device=dev).repeat(B, 1).requires_grad_(True)
        s = torch.ones(B, device=dev, requires_grad=True)
        fused = chore_b200.FusedFitSteps(net, split, data, R, t, s, fitter=fit, phase="kpts")
        fused.data = data
        return split, (R, t, s), fused

    return fit, cc, build_state


def bench_gen(net, dev, img_host, args):
    """Generator.gen_pc_batch for 'human' and 'object' (recon/generator.py:96-121): init 30 k samples, then outer
    iterations of {10 projection steps (query df + gradient to the points), surface filter, resample 20 k}.  With random
    weights the field is not a distance field, so the 4 mm filter is opened up (every sample passes) and the target
    count fixes the number of outer iterations at 4 per field = 80 forward + 80 backward queries of 20-30 k points."""
    import chore_b200
    B = img_host.shape[0]
    gen = chore_b200.Generator(net, filter_val=1e9, device=str(dev), rng=args.rng)
    cc = torch.tensor([[1008., 995.]], device=dev).repeat(B, 1)
    net.filter(img_host.to(dev))

    def job():
        out = {}
        for t in ("human", "object"):
            init = gen.init_samples(30000, batch_size=B)
            out[t] = gen.gen_pc_batch(net, t, init, 3 * args.points, {"crop_center": cc}, num_steps=10, sample_num=args.points)
        return out

    times = []
    for rep in range(args.reps + 1):
        torch.manual_seed(rep)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = job()
        b.record()
        torch.cuda.synchronize()
        if rep > 0:
            times.append(a.elapsed_time(b))
    n = out["human"]["points"].shape[1]
    ms = sorted(times)[len(times) // 2]
    print(json.dumps({"metric": "gen_pc_batch_ms", "value": ms, "unit": "ms per image batch (human + object fields)", "batch": B,
                      "points_returned_per_field": n, "outer_iterations_per_field": 4, "projection_steps": 10,
                      "queries": "80 x (df forward + gradient to the points) of 20-30k points + 8 x all-head forward",
                      "higher_is_better": False, "data": "synthetic", "dtype": "f32",
                      "rng": args.rng + (" (resampling on the device, no host round trip inside an outer iteration)" if args.rng == "device"
                                         else " (torch CPU draws in the reference's order: bit-identical sample sets)"),
                      "note": "reference on CPU: 1.44 s per forward+backward at 20 k points (SURVEY.md section 6) => ~2 min for the same loop"}),
          flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1, help="images in total (sharded over the ranks)")
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--points", type=int, default=20000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true", help="launch the fused steps directly (for ncu launch lists)")
    ap.add_argument("--rng", default="device", choices=["device", "reference"], help="--gen: where the resampling draws come from")
    ap.add_argument("--gen", action="store_true", help="time Generator.gen_pc_batch (neural point cloud stage) instead of the fit loop")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench_fit.py needs a CUDA device (no CPU fallback for the product path)")
    import torch.distributed as dist
    import chore_b200
    from chore_b200 import dist as cdist
    from oracle import chore_oracle as O          # input / weight synthesis only
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mine = cdist.shard_images(args.batch, rank, world)
    B = len(mine)
    assert B > 0, "more ranks than images"
    net = chore_b200.CHORE(device=str(dev))
    net.load_state_dict(O.make_state_dict(0, "unit"))
    img_host = torch.cat([O.synth_images(1000 + i, B=1, size=512) for i in mine]).pin_memory()
    pts = O.synth_points("init_box", 5 + rank, B, args.points).to(dev)
    if args.gen:
        return bench_gen(net, dev, img_host, args)
    fit, cc, build_state = make_fit_problem(net, dev, B, args.points, seed=100 + rank)

    def build():
        split, (R, t, s), fused = build_state()
        # product path: SMPL step + object step forked onto two streams inside ONE CUDA graph per iteration
        return split, (R, t, s), ((fused.smpl_step, fused.object_step) if args.no_graph else (fused.graphed_iteration(), None)), fused

    def job(state):
        split, (R, t, s), (g_smpl, g_obj), fused = state
        net.filter(img_host.to(dev, non_blocking=True))
        net.query(pts, crop_center=cc)
        for it in range(args.iters):
            if it % 10 == 0:
                fused.zero_grad()                    # the reference loops: optimizer.zero_grad() once per 10 inner steps
            g_smpl()
            if g_obj is not None:
                g_obj()
        fitted = torch.cat([split.global_pose, split.body_pose, split.hand_pose, split.top_betas, split.other_betas, split.trans,
                            R.reshape(B, 9), t, s.reshape(B, 1)], 1).detach()
        return cdist.gather_results(fitted.t().contiguous(), dim=1).t()      # (batch, 182) on every rank

    net.filter(img_host.to(dev))                     # features must exist before the graphs are captured
    launches0 = chore_b200.launch_count()
    times = []
    for rep in range(args.reps + 1):                 # first repetition = warm-up (graph capture, allocator)
        state = build()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = job(state)
        b.record()
        torch.cuda.synchronize()
        if rep > 0:
            times.append(a.elapsed_time(b))
    assert out.shape == (args.batch, 182) and torch.isfinite(out).all()
    ms = torch.tensor([sorted(times)[len(times) // 2]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = ms.item() / 1e3
        print(json.dumps({
            "metric": "fit_iters_per_sec", "value": args.batch * args.iters / sec, "unit": "image-iterations/s",
            "n_gpus": world, "batch": args.batch, "images_per_gpu": B, "iters": args.iters, "points_per_image": args.points,
            "ms_per_job": ms.item(), "ms_per_iteration": ms.item() / args.iters, "reps": args.reps,
            "queries_points_per_iteration_per_image": 6890 + args.points,
            "config": {"workload": f"{args.batch} x 5x512x512 images -> encoder -> {args.points}-point query -> {args.iters} fit iterations "
                                   "(SMPL-H step with every forward_smpl term + object-only step), fitted parameters all-gathered",
                       "parallelism": f"batch axis over {world} rank(s), NCCL all-gather of 182 floats per image"},
            "data": "synthetic", "dtype": "f32"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

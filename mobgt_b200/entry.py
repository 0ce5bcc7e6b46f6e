"""`entry.py` of the reference (graphormer/entry.py:48-165) with the same flag surface, driving the B200-native hot path.

    python -m mobgt_b200.entry --dataset_name toyotagraph --gpus 1 --batch_size 256 --hidden_dim 128 --num_heads 8 \
        --n_layers 6 --ffn_dim 1024 --dropout_rate 0.1 --peak_lr 2e-4 --edge_type multi_hop --multi_hop_max_dist 20 \
        --warmup_updates 40000 --tot_updates 400000 --seed 1 --max_epochs 1 --default_root_dir exps/toyota
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m mobgt_b200.entry --accelerator ddp ...

Flag groups (names and defaults as in the reference):
  * model:      Graphormer.add_model_specific_args                      (model_fqandtoyo.py:1619-1641)
  * datamodule: --dataset_name --num_workers --batch_size --seed --multi_hop_max_dist --rel_pos_max   (data.py:197-207)
  * trainer:    the pytorch_lightning.Trainer flags the README uses      (README.md:62): --gpus --accelerator --precision
                --max_epochs --max_steps --check_val_every_n_epoch --default_root_dir --resume_from_checkpoint
The Lightning runtime itself is out of scope: a minimal torch loop replaces it — data-parallel over
trajectory graphs with one NCCL all-reduce of a flat fp32 gradient buffer per step, `last.ckpt` auto-resume
(entry.py:135-137) and the reference's three metric lines at test time (model_fqandtoyo.py:1593-1595).

The reference reads ../dataset/<name>/raw/*; those blobs are not shipped, so the data here is the
seeded synthetic world of mobgt_b200.synth (`--synthetic` picks the BASELINE.json shape, `--train_graphs` the split size).
"""
import argparse
import os
import time

import numpy as np
import torch

DATASET_DEFAULT_SHAPE = {"foursquaregraph": "c1", "toyotagraph": "c2", "gowalla_nevda": "c4", "gowalla_7day": "c4"}
COLLATORS = {"foursquaregraph": "collator_foursquare", "toyotagraph": "collator_toyota", "gowalla_nevda": "collator_gowalla",
             "gowalla_7day": "collator_gowalla"}


def build_parser():
    from .model import Graphormer
    parser = argparse.ArgumentParser(description="MobGT (B200-native hot path)")
    t = parser.add_argument_group("pl.Trainer")                         # entry.py:53
    t.add_argument("--gpus", type=int, default=1)
    t.add_argument("--accelerator", type=str, default=None)             # "ddp" -> one rank per GPU (launched by torchrun)
    t.add_argument("--precision", type=int, default=16)                 # 16 -> bf16 GEMMs/attention (the reference: fp16 AMP)
    t.add_argument("--max_epochs", type=int, default=1)
    t.add_argument("--max_steps", type=int, default=None)               # overwritten from tot_updates (entry.py:57)
    t.add_argument("--check_val_every_n_epoch", type=int, default=1)
    t.add_argument("--default_root_dir", type=str, default="exps/mobgt")
    t.add_argument("--resume_from_checkpoint", type=str, default=None)
    t.add_argument("--progress_bar_refresh_rate", type=int, default=50)
    Graphormer.add_model_specific_args(parser)                          # entry.py:54
    d = parser.add_argument_group("GraphDataModule")                    # entry.py:55 / data.py:197-207
    d.add_argument("--dataset_name", type=str, default="toyotagraph")
    d.add_argument("--num_workers", type=int, default=0)
    d.add_argument("--batch_size", type=int, default=256)
    d.add_argument("--seed", type=int, default=1)
    d.add_argument("--multi_hop_max_dist", type=int, default=5)
    d.add_argument("--rel_pos_max", type=int, default=1024)
    s = parser.add_argument_group("synthetic data (stand-in for ../dataset/<name>)")
    s.add_argument("--synthetic", type=str, default=None, help="c1|c2|c4|tiny (default: by dataset_name)")
    s.add_argument("--train_graphs", type=int, default=2048)
    s.add_argument("--test_graphs", type=int, default=512)
    return parser


def parse_args(argv=None):
    args = build_parser().parse_args(argv)
    args.max_steps = args.tot_updates + 1                               # entry.py:57
    return args


def _ckpt_dir(args):
    return os.path.join(args.default_root_dir, "lightning_logs", "checkpoints")     # entry.py:123


def cli_main(argv=None):
    import torch.distributed as dist
    from . import _C, collator, synth
    from .model import Graphormer
    args = parse_args(argv)
    rank, world_size = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _C.require_cuda()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0 and not args.test and not args.validate:
        print(args)
    torch.manual_seed(args.seed)                                        # pl.seed_everything (entry.py:60)
    np.random.seed(args.seed)

    shape = args.synthetic or DATASET_DEFAULT_SHAPE.get(args.dataset_name, "c2")
    world = synth.make_world(shape, seed=args.seed, dataset_name=args.dataset_name)
    cap = synth.CONFIGS[shape]["cap"]
    train = synth.make_items(world, args.train_graphs, cap, seed=args.seed, cfg_id=11)
    test = synth.make_items(world, args.test_graphs, cap, seed=args.seed, cfg_id=12)
    collate = getattr(collator, COLLATORS[args.dataset_name])
    latlon = torch.from_numpy(world.latlon).to(dev)

    def batches(items, shuffle, epoch):
        order = np.arange(len(items))
        if shuffle:
            np.random.default_rng([args.seed, epoch]).shuffle(order)
        order = order[rank::world_size]                                 # DistributedSampler: graphs g = r (mod world)
        for i in range(0, len(order), args.batch_size):
            sel = [items[j] for j in order[i:i + args.batch_size]]
            yield collate(sel, max_node=512, multi_hop_max_dist=args.multi_hop_max_dist, rel_pos_max=args.rel_pos_max,
                          world=world, latlon_dev=latlon, device=dev)

    model = Graphormer(
        n_layers=args.n_layers, num_heads=args.num_heads, hidden_dim=args.hidden_dim,
        attention_dropout_rate=args.attention_dropout_rate, dropout_rate=args.dropout_rate,
        intput_dropout_rate=args.intput_dropout_rate, weight_decay=args.weight_decay, ffn_dim=args.ffn_dim,
        dataset_name=args.dataset_name, warmup_updates=args.warmup_updates, tot_updates=args.tot_updates, peak_lr=args.peak_lr,
        end_lr=args.end_lr, edge_type=args.edge_type, multi_hop_max_dist=args.multi_hop_max_dist, flag=args.flag,
        flag_m=args.flag_m, flag_step_size=args.flag_step_size, world=world).to(dev)
    if args.checkpoint_path != "":                                      # entry.py:71-93 (strict=False)
        sd = torch.load(args.checkpoint_path, map_location=dev)
        model.load_state_dict(sd.get("state_dict", sd), strict=False)
    if rank == 0:
        print("total params:", sum(p.numel() for p in model.parameters()))

    (opt,), (sched_cfg,) = model.configure_optimizers()
    sched = sched_cfg["scheduler"]
    params = list(model.parameters())
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()

    ckdir = _ckpt_dir(args)
    last = os.path.join(ckdir, "last.ckpt")
    step, epoch0 = 0, 0
    resume = args.resume_from_checkpoint
    if not args.test and not args.validate and os.path.exists(last):    # entry.py:135-137
        resume = last
    if resume:
        ck = torch.load(resume, map_location=dev)
        model.load_state_dict(ck["state_dict"], strict=False)
        if "optimizer" in ck and not (args.test or args.validate):
            opt.load_state_dict(ck["optimizer"])
            sched.load_state_dict(ck["lr_scheduler"])
            step, epoch0 = ck.get("global_step", 0), ck.get("epoch", 0)
        if rank == 0:
            print("args.resume_from_checkpoint", resume)

    def evaluate(items, tag):
        model.eval()
        outs = []
        with torch.no_grad():
            for b in batches(items, False, 0):
                outs.append(model.test_step(b))
        res = model.test_epoch_end(outs) if rank == 0 or world_size == 1 else None
        model.train()
        return res

    if args.test or args.validate:
        print(evaluate(test, "test"))
    else:
        model.train()
        t0 = time.time()
        for epoch in range(epoch0, args.max_epochs):
            for b in batches(train, True, epoch):
                if step >= args.max_steps:
                    break
                flat.zero_()
                loss = model.training_step(b)
                loss.backward()
                if world_size > 1:
                    dist.all_reduce(flat)
                    flat.div_(world_size)
                opt.step()
                sched.step()
                step += 1
                if rank == 0 and step % max(1, args.progress_bar_refresh_rate) == 0:
                    print(f"epoch {epoch} step {step} train_loss {loss.item():.5f} lr {sched.get_last_lr()[0]:.3e} "
                          f"({time.time() - t0:.1f}s)")
            if (epoch + 1) % args.check_val_every_n_epoch == 0:
                evaluate(test, "valid")
            if rank == 0:
                os.makedirs(ckdir, exist_ok=True)
                torch.save({"state_dict": model.state_dict(), "optimizer": opt.state_dict(), "lr_scheduler": sched.state_dict(),
                            "global_step": step, "epoch": epoch + 1, "hyper_parameters": vars(args)}, last)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    cli_main()

"""CUDA-graph capture of the training step's forward + backward (the launch-bound inner loop of the path).

A MobGT step is ~550 kernel launches of a few microseconds each; enqueuing them from Python costs as much host time as the
GPU needs to run them.  When consecutive batches have the SAME packed shapes (fixed node count per graph, e.g. every
trajectory at the cap) the launches are captured once and replayed:

    g = GraphedTrainStep(model, flat_grads, example_batch)      # warm-up + capture; raises GraphCaptureError if not capturable
    g.load(batch)         # copy a freshly collated batch into the static buffers (raises ShapeMismatch -> run it eagerly)
    loss = g.run()        # zero grads, forward, loss, backward — one cudaGraphLaunch
    # gradient all-reduce (NCCL) and the optimizer step stay outside the graph

Everything the captured kernels read lives in static memory: the batch fields, the sort plans of the K4 backward, the bf16
working copies of the weights (re-cast from the fp32 masters inside the graph), and a u64 step counter in device memory that
the fused dropout kernels fold into their seed (incremented inside the graph, so every replay draws fresh masks; torch's own
dropouts use torch's graph-safe Philox state).  Batches of other shapes (the natural node-count distribution) are not
capturable this way and run eagerly.
"""
import torch

from . import _C, ops


class GraphCaptureError(RuntimeError):
    pass


class ShapeMismatch(ValueError):
    pass


_PLAN_KEYS = ("poi", "slot", "pos", "ind", "outd")


def _tensor_fields(batch):
    return {k: v for k, v in batch.__dict__.items() if isinstance(v, torch.Tensor) and not k.startswith("_")}


class GraphedTrainStep:
    def __init__(self, model, flat_grads, example_batch, warmup=3, after_backward=None, pool=None, split_head=False):
        """after_backward: optional callable run right after loss.backward() inside the captured region (the trainer joins
        the side stream of its early gradient all-reduce there, so the collective is part of the graph).
        split_head: capture TWO graphs — A = zero grads + forward + loss + the backward of the two heads (which completes the
        gradient of out_proj), B = the encoder backward — so that the data-parallel trainer can start the all-reduce of the
        out_proj bucket between them (`run_a()`, collective on a side stream, `run_b()`)."""
        self.model, self.flat = model, flat_grads
        self._after_backward = after_backward
        self.split = bool(split_head)
        self.graph_b = torch.cuda.CUDAGraph() if self.split else None
        self.static = example_batch
        self.dev = flat_grads.device
        if "_plans" not in example_batch.__dict__:
            example_batch.build_plans()
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._fields = _tensor_fields(example_batch)
        self._meta = (int(example_batch.B), int(example_batch.N), {k: tuple(v.shape) for k, v in self._fields.items()})
        self._all_large = ops.all_large(example_batch)     # decides which attention kernels the capture contains
        self.graph = torch.cuda.CUDAGraph()
        ops.set_device_seed(self.seed_dev)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if hasattr(model, "_w16"):
                model._w16.stamp = None          # the bf16 re-cast of the weights must be part of the captured work
            n0 = _C.launch_count()
            if self.split:
                if pool is None:
                    pool = torch.cuda.graph_pool_handle()          # A's activations are read by B: one pool for both
                with torch.cuda.graph(self.graph, pool=pool):
                    self.loss, z, z_cut = self._body_a()
                with torch.cuda.graph(self.graph_b, pool=pool):
                    self._body_b(z, z_cut)
                del z, z_cut
            else:
                with torch.cuda.graph(self.graph, pool=pool):      # pool: memory shared by graphs that never run concurrently
                    self.loss = self._body()
            self.launches = _C.launch_count() - n0       # libmobgt kernel nodes in the graph(s) = launches per replay
            torch.cuda.synchronize()
        except Exception as e:                   # noqa: BLE001 — anything that is illegal during capture: report, let the caller fall back
            ops.set_device_seed(None)
            raise GraphCaptureError(f"training step is not capturable: {type(e).__name__}: {e}") from e
        finally:
            # eager steps (other shapes) must not use the graph's counter
            ops.set_device_seed(None)

    def _body_a(self):
        self.seed_dev.add_(1)
        self.flat.zero_()
        ops.grads_zeroed(self.flat)
        loss, z, z_cut = self.model.training_step(self.static, split=True)
        loss.backward()                       # heads only: the graph is cut at z
        return loss, z, z_cut

    def _body_b(self, z, z_cut):
        z.backward(z_cut.grad)                # encoder, embeddings, bias tables, GCNs
        if self._after_backward is not None:
            self._after_backward()

    def _body(self):
        if self.split:
            loss, z, z_cut = self._body_a()
            self._body_b(z, z_cut)
            return loss
        self.seed_dev.add_(1)
        self.flat.zero_()
        ops.grads_zeroed(self.flat)
        loss = self.model.training_step(self.static)
        loss.backward()
        if self._after_backward is not None:
            self._after_backward()
        return loss

    def load(self, batch):
        """Copy a collated batch into the static buffers (stream-ordered: safe right after the previous replay)."""
        if batch is self.static:
            return
        f = _tensor_fields(batch)
        if (int(batch.B), int(batch.N)) != self._meta[:2] or {k: tuple(v.shape) for k, v in f.items()} != self._meta[2] \
                or ops.all_large(batch) != self._all_large:
            raise ShapeMismatch("batch shapes differ from the captured ones")
        for k, dst in self._fields.items():
            dst.copy_(f[k], non_blocking=True)
        src_plans = batch.__dict__.get("_plans") or batch.build_plans()
        dst_plans = self.static.__dict__["_plans"]
        for key in _PLAN_KEYS:
            for d, s in zip(dst_plans[key], src_plans[key]):
                d.copy_(s, non_blocking=True)
        dst_plans["node_rows"].copy_(src_plans["node_rows"], non_blocking=True)
        if "cat" in dst_plans:                   # built lazily by EmbedGather.backward from the model's cat_of_poi table
            cat = ops.sort_plan(self.model.cat_of_poi[batch.x_nodes.long() - 1].long() - 1)
            for d, s in zip(dst_plans["cat"], cat):
                d.copy_(s, non_blocking=True)
        self.static.n_host = batch.n_host
        self.static.padded = getattr(batch, "padded", False)

    def run(self):
        self.graph.replay()
        if self.split:
            self.graph_b.replay()
        return self.loss

    def run_a(self):
        """split_head: zero + forward + loss + head backward; out_proj's gradient is final when this replay has finished."""
        self.graph.replay()
        return self.loss

    def run_b(self):
        self.graph_b.replay()

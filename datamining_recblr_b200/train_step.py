"""CUDA-graph form of RecBole's `Trainer._train_epoch` body for one batch:

    optimizer.zero_grad(); loss = model.calculate_loss(interaction); loss.backward(); optimizer.step()

At RecBLR's shapes (L = 50..200, D = 64) a step is ~250 small launches and host-launch-bound when run eagerly; captured
once and replayed it is bound by the kernels.  Everything the step touches is graph-safe: the libbdlru kernels take
no host syncs and allocate nothing, the fused front end's dropout stream advances through a device-side counter, and
the optimizer must be constructed with `capturable=True`.  The caller keeps RecBole's contract: pass an interaction
(dict of tensors) per step, get the loss tensor back (read it with `.item()` only when the trainer needs it).
"""
import torch


class GraphedTrainStep:
    def __init__(self, model, optimizer, example_interaction, autocast_dtype=None, grad_hook=None, warmup=3):
        """example_interaction: dict of CUDA tensors with the shapes/dtypes of every later batch.
        grad_hook(params): optional callable run between backward and optimizer.step (e.g. the data-parallel
        gradient all-reduce); it is captured too."""
        self.model, self.optimizer = model, optimizer
        self.autocast_dtype, self.grad_hook = autocast_dtype, grad_hook
        self.static = {k: v.clone() for k, v in example_interaction.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up on a side stream (allocator pools, workspaces, cuBLAS handles)
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()

    def _body(self):
        self.optimizer.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=self.autocast_dtype or torch.bfloat16, enabled=self.autocast_dtype is not None):
            loss = self.model.calculate_loss(self.static)
        loss.backward()
        if self.grad_hook is not None:
            self.grad_hook([p for p in self.model.parameters() if p.grad is not None])
        self.optimizer.step()
        return loss.detach()

    def __call__(self, interaction):
        """Replays the captured step when the batch has the captured shapes; any other batch (the ragged last batch of
        a RecBole epoch) runs the same body eagerly — same kernels, same optimizer state, no graph."""
        if any(tuple(interaction[k].shape) != tuple(v.shape) for k, v in self.static.items()):
            return self._eager(interaction)
        for k, v in self.static.items():
            v.copy_(interaction[k], non_blocking=True)
        self.graph.replay()
        return self.loss

    def _eager(self, interaction):
        static, self.static = self.static, {k: interaction[k] for k in self.static}
        try:
            return self._body()
        finally:
            self.static = static

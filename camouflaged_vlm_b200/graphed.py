"""CUDA-graph replay of the encoder for a fixed batch size (latency path, e.g. the reference's batch-1 demo).

One forward is 243 dependent kernel launches; replayed from a graph the gaps between them shrink (measured on a B200,
tools/graph_probe.py: batch 1: 7.84 -> 7.42 ms, batch 8: 47.8 -> 46.3 ms, bit-identical outputs).  A graph bakes its
buffers in, so this is an explicit wrapper with static-buffer semantics rather than something ``forward`` does silently:
the tensors it returns are overwritten by the next call."""
from typing import List, Tuple, Union

import torch

from .image_encoder import ImageEncoderViT


def _pin_runtime_buffers(module):
    """References to every device buffer a captured forward of ``module`` has baked in (workspace, packed weights), plus
    the key of the parameters they were packed from."""
    dev = next(module.parameters()).device
    return {"workspace": getattr(module, "_workspace", None), "packed": getattr(module, "_packed", None),
            "device": dev, "weights_key": module._weights_key(dev)}


def _owners_of(fn, extra):
    """The modules of this package that ``fn`` runs: explicit ``owners``, the object of a bound method / partial, and
    whatever a lambda or nested function closes over."""
    found = list(extra or [])

    def add(obj):
        if hasattr(obj, "_weights_key") and hasattr(obj, "_packed_key") and all(obj is not f for f in found):
            found.append(obj)

    seen = fn
    while seen is not None:
        add(getattr(seen, "__self__", None))
        for cell in getattr(seen, "__closure__", None) or ():
            try:
                add(cell.cell_contents)
            except ValueError:          # empty cell
                pass
        seen = getattr(seen, "func", None)          # functools.partial
    return found


class GraphedEncoder:
    def __init__(self, model: ImageEncoderViT, batch: int, interm: bool = False):
        cfg = model.config
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedEncoder needs the model on a CUDA device")
        self.model, self.batch, self.interm = model, batch, interm
        self.static_x = torch.zeros(batch, cfg.in_chans, cfg.img_size, cfg.img_size, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():          # warm-up outside the capture: weight packing,
            model(self.static_x, interm=interm)                 # workspace, cudaFuncSetAttribute, tensor-map cache
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = model(self.static_x, interm=interm)
        self._weights_key = model._packed_key
        # The graph has the module's workspace and packed bf16 weights baked in as raw pointers.  Both were allocated in
        # the warm-up (outside the graph's pool) and the module replaces them when a later eager call needs a bigger
        # workspace or the parameters change: hold references so they outlive the module's, and check at replay.
        self._keep = _pin_runtime_buffers(model)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> Union[torch.Tensor, Tuple[torch.Tensor, List[torch.Tensor]]]:
        """``x``: [batch, 3, S, S] on the model's device.  Returns the graph's static output tensors (valid until the
        next call)."""
        if tuple(x.shape) != tuple(self.static_x.shape):
            raise ValueError(f"captured for input {tuple(self.static_x.shape)}, got {tuple(x.shape)}")
        if self.model._weights_key(self.static_x.device) != self._weights_key:
            raise RuntimeError("the model's parameters changed after capture: build a new GraphedEncoder")
        # (a replaced workspace is harmless: self._keep keeps the captured one alive and the graph keeps using it)
        self.static_x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out


class GraphedCall:
    """CUDA-graph replay of any of this package's forward calls for fixed shapes (the mask decoder's 37 launches, the
    CLIP tower's 183): ``GraphedCall(fn, *example_tensors)`` runs ``fn(*static_copies)`` once to warm up, captures it, and
    ``__call__(*tensors)`` copies the arguments into the static inputs and replays.  Non-tensor arguments are baked in.
    Outputs are the graph's static tensors, overwritten by the next call; parameters must not change after capture
    (checked at replay for the modules ``fn`` is known to run)."""

    def __init__(self, fn, *example, owners=None):
        """``owners``: the modules whose forward ``fn`` runs when that cannot be seen from ``fn`` itself (a lambda); a
        bound method or functools.partial of a module is recognised without it."""
        tensors = [a for a in example if isinstance(a, torch.Tensor)]
        if not tensors or any(not t.is_cuda for t in tensors):
            raise RuntimeError("GraphedCall needs CUDA tensor arguments")
        dev = tensors[0].device
        self.fn = fn
        self.static_in = [a.clone() if isinstance(a, torch.Tensor) else a for a in example]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            fn(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = fn(*self.static_in)
        # same reasoning as GraphedEncoder: keep the baked-in workspace / packed weights alive, detect re-packing
        self._owners = _owners_of(fn, owners)
        self._keep = [_pin_runtime_buffers(m) for m in self._owners]

    @torch.no_grad()
    def __call__(self, *args):
        if len(args) != len(self.static_in):
            raise ValueError(f"captured with {len(self.static_in)} arguments, got {len(args)}")
        for m, keep in zip(self._owners, self._keep):
            if m._weights_key(keep["device"]) != keep["weights_key"]:
                raise RuntimeError(f"{type(m).__name__}: parameters changed after capture (updated, moved or "
                                   "re-loaded): build a new GraphedCall")
        for s, a in zip(self.static_in, args):
            if isinstance(s, torch.Tensor):
                if tuple(s.shape) != tuple(a.shape) or s.dtype != a.dtype:
                    raise ValueError(f"captured for {tuple(s.shape)} {s.dtype}, got {tuple(a.shape)} {a.dtype}")
                s.copy_(a, non_blocking=True)
        self.graph.replay()
        return self.static_out

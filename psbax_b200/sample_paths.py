"""GP posterior sample paths on the GPU: drawing their parameters and the fused
evaluate-and-reduce operations, over the C ABI of ``libpsbax_b200.so``.

Replaces, for the hot path, what ``get_function_samples`` builds from botorch
(``/root/reference/src/utils.py:185-250``): instead of one opaque callable per
sample evaluated in Python chunks of 100 candidates
(``src/algorithms/topk.py:48-49``), a ``SamplePathBatch`` holds the parameters
of S paths in HBM and evaluates all of them over X* in one kernel launch,
reduction included.

    f_s(x) = y_mean + y_std * ( mean_const + sum_l W[s,l] cos(omega[g,l].x + b[g,l])
                                + sum_j V[s,j] kappa(|x/ls - Xtrain_j|) )

There is no CPU implementation here: every evaluation goes through the CUDA
library and raises ``PsbaxError`` if it is unavailable.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ENGINE_IDS, KERNEL_IDS, MeshStruct, PathsStruct, PsbaxError

F32 = torch.float32
F64 = torch.float64


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise PsbaxError("no CUDA device: psbax_b200 evaluates sample paths on the GPU only")
    return torch.device("cuda", torch.cuda.current_device())


@dataclass
class LevelSetResult:
    """Per-sample level-set reduction (``src/algorithms/levelset.py:34-39``)."""

    mask: Optional[torch.Tensor]  # int32 words [S, ceil(N/32)], bit i%32 of word i//32
    count: torch.Tensor  # int64 [S]
    maxval: torch.Tensor  # float32 [S]
    argmax: torch.Tensor  # int64 [S] (row_offset + row)
    N: int
    pooled: Optional[torch.Tensor] = None  # int32 words [ceil(N/32)]: OR of the S masks (computed in the kernel)

    def indices(self, s: int) -> torch.Tensor:
        """Row indices of sample ``s``'s super-level set, or ``[argmax]`` when empty."""
        if int(self.count[s]) == 0:
            return self.argmax[s:s + 1].clone()
        words = self.mask[s]
        bits = (words.unsqueeze(1) >> torch.arange(32, device=words.device, dtype=torch.int32)) & 1
        return torch.nonzero(bits.reshape(-1)[: self.N], as_tuple=False).reshape(-1)


def _stream_bounds(N: int, c: int):
    """Chunk boundaries of the streamed level set: chunks of ``c`` rows (a multiple of 2048) with short
    chunks (c/8, c/2) at both ends — the first upload and the last download are the only copies
    that do not overlap with a kernel.  Every boundary but the last is a multiple of 256."""
    if N < 4 * c:
        return list(range(0, N, c)) + [N]
    sizes = [c // 8, c // 2]
    left = N - sum(sizes)
    tail = c // 2 + c // 8
    while left - tail > c:
        sizes.append(c)
        left -= c
    mid = ((left - tail) // 256) * 256
    if mid > 0:
        sizes.append(mid)
        left -= mid
    sizes.append(c // 2)
    left -= c // 2
    sizes.append(left)  # c/8 .. c/8 + 255 rows
    bounds = [0]
    for sz in sizes:
        bounds.append(bounds[-1] + sz)
    return bounds


class DeviceFFNN:
    """The deep-kernel GP's feed-forward embedding (``src/models/deep_kernel_gp.py:18-84``) as a CUDA
    op: a float32 device COPY of the network's Linear layers (the caller's module is not moved or
    modified) evaluated by ``psbax_mlp_forward`` — X [N, d_in] -> [N, h] float32 on the GPU."""

    def __init__(self, net, device=None):
        self.device = torch.device(device) if device is not None else default_device()
        self.lib = _lib.load()
        linears = [m for m in net.modules() if isinstance(m, torch.nn.Linear)]
        if not 1 <= len(linears) <= _lib.MLP_MAX_LAYERS:
            raise PsbaxError(f"embedding network with {len(linears)} Linear layers (supported: 1..{_lib.MLP_MAX_LAYERS})")
        act = str(getattr(net, "activation", "leaky_relu")).lower()
        if act not in _lib.ACTIVATION_IDS:
            raise PsbaxError(f"activation {act!r} has no CUDA form")
        with torch.no_grad():
            self.W = [m.weight.detach().to(self.device, F32).T.contiguous() for m in linears]  # [in, out]
            self.b = [(m.bias.detach() if m.bias is not None else torch.zeros(m.out_features)).to(self.device, F32).contiguous()
                      for m in linears]
        self.widths = [self.W[0].shape[0]] + [w.shape[1] for w in self.W]
        if max(self.widths) > 256:
            raise PsbaxError("embedding layers wider than 256 are not supported by psbax_mlp_forward")
        self.h = self.widths[-1]
        st = _lib.MlpStruct(n_layers=len(linears), activation=_lib.ACTIVATION_IDS[act], reserved=0,
                            negative_slope=0.01, reserved_f=0.0)
        for i, w in enumerate(self.widths):
            st.width[i] = w
        for i, (w, b) in enumerate(zip(self.W, self.b)):
            st.W[i], st.b[i] = w.data_ptr(), b.data_ptr()
        self._struct = st

    def __call__(self, X) -> torch.Tensor:
        X = torch.as_tensor(X)
        X = X.reshape(-1, X.shape[-1])
        if X.shape[1] != self.widths[0]:
            raise PsbaxError(f"embedding expects {self.widths[0]} input columns, got {X.shape[1]}")
        if X.device != self.device or X.dtype != F32 or not X.is_contiguous():
            X = X.to(self.device, F32).contiguous()
        with torch.cuda.device(self.device):
            out = torch.empty(X.shape[0], self.h, dtype=F32, device=self.device)
            if X.shape[0]:
                _lib.check(self.lib.psbax_mlp_forward(C.byref(self._struct), X.data_ptr(), X.shape[0], out.data_ptr(),
                                                      torch.cuda.current_stream(self.device).cuda_stream),
                           "psbax_mlp_forward")
        return out


class SamplePathBatch:
    """S sample paths resident on one GPU."""

    def __init__(self, omega, bias, W, kernel: str = "matern52", Xtrain=None, inv_ls=None, V=None,
                 mean_const: float = 0.0, y_mean: float = 0.0, y_std: float = 1.0,
                 device: Optional[torch.device] = None, engine: str = "auto", input_transform=None):
        """``input_transform``: optional callable applied to the candidates on the GPU before the
        evaluation (the DKGP embedding, ``src/utils.py:202``); the paths then live on its output
        space (``d`` = embedding width)."""
        self.input_transform = input_transform
        self.device = torch.device(device) if device is not None else default_device()
        if self.device.type != "cuda":
            raise PsbaxError("SamplePathBatch needs a CUDA device")
        self.lib = _lib.load()

        def dev(t):
            return None if t is None else torch.as_tensor(t).to(self.device, F32).contiguous()

        self.omega, self.bias, self.W = dev(omega), dev(bias), dev(W)
        self.Xtrain, self.inv_ls, self.V = dev(Xtrain), dev(inv_ls), dev(V)
        if self.omega.dim() != 3 or self.W.dim() != 2:
            raise PsbaxError("omega must be [G, L, d] and W [S, L]")
        self.G, self.L, self.d = self.omega.shape
        self.S = self.W.shape[0]
        self.n = 0 if self.V is None else self.V.shape[1]
        if self.n == 0:
            self.Xtrain = self.inv_ls = self.V = None
        self.kernel = kernel
        self.engine = engine
        self.mean_const, self.y_mean, self.y_std = float(mean_const), float(y_mean), float(y_std)
        if kernel not in KERNEL_IDS:
            raise NotImplementedError(f"kernel {kernel!r}")  # botorch RandomFourierFeatures raises too
        if tuple(self.bias.shape) != (self.G, self.L) or self.W.shape[1] != self.L:
            raise PsbaxError("bias must be [G, L] and W [S, L]")
        if self.n and (tuple(self.Xtrain.shape) != (self.n, self.d) or self.V.shape[0] != self.S
                       or self.inv_ls.numel() != self.d):
            raise PsbaxError("Xtrain must be [n, d], V [S, n], inv_ls [d]")
        self._struct = PathsStruct(
            d=self.d, L=self.L, G=self.G, S=self.S, n=self.n, kernel=KERNEL_IDS[kernel],
            engine=ENGINE_IDS[engine], reserved=0,
            omega=_ptr(self.omega) if self.L else None, bias=_ptr(self.bias) if self.L else None,
            W=_ptr(self.W) if self.L else None, Xtrain=_ptr(self.Xtrain), inv_ls=_ptr(self.inv_ls),
            V=_ptr(self.V), mean_const=self.mean_const, y_mean=self.y_mean, y_std=self.y_std,
            reserved_f=0.0)
        self._pack = None
        self._ws = None
        self._build_pack()

    # -- plumbing --------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _build_pack(self) -> None:
        nbytes = C.c_size_t(0)
        _lib.check(self.lib.psbax_pack_bytes(C.byref(self._struct), C.byref(nbytes)), "psbax_pack_bytes")
        with torch.cuda.device(self.device):
            self._pack = torch.empty((nbytes.value + 3) // 4, dtype=F32, device=self.device)
            _lib.check(self.lib.psbax_pack(C.byref(self._struct), self._pack.data_ptr(),
                                           self._pack.numel() * 4, self._stream()), "psbax_pack")

    def _workspace(self, N: int, k: int) -> torch.Tensor:
        nbytes = C.c_size_t(0)
        _lib.check(self.lib.psbax_workspace_bytes(C.byref(self._struct), N, k, C.byref(nbytes)),
                   "psbax_workspace_bytes")
        if self._ws is None or self._ws.numel() < nbytes.value:
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        return self._ws

    def _x(self, X) -> torch.Tensor:
        X = torch.as_tensor(X)
        if self.input_transform is not None:
            with torch.no_grad():
                X = self.input_transform(X.reshape(-1, X.shape[-1]).to(self.device))
        if X.shape[-1] != self.d:
            raise PsbaxError(f"X has last dim {X.shape[-1]}, paths expect d={self.d}")
        X = X.reshape(-1, self.d)
        if X.device != self.device or X.dtype != F32 or not X.is_contiguous():
            X = X.to(self.device, F32).contiguous()
        return X

    # -- operations ------------------------------------------------------------------
    def values(self, X) -> torch.Tensor:
        """f_s(X[i]) for every sample: [N, S] float32 on the GPU."""
        X = self._x(X)
        N = X.shape[0]
        with torch.cuda.device(self.device):
            out = torch.empty(N, self.S, dtype=F32, device=self.device)
            _lib.check(self.lib.psbax_eval_values(C.byref(self._struct), self._pack.data_ptr(),
                                                  X.data_ptr(), N, out.data_ptr(), self.S,
                                                  self._stream()), "psbax_eval_values")
        return out

    def sumsq(self, X) -> torch.Tensor:
        """``sum_s f_s(X[i])**2`` for every row: [N] float32 on the GPU (the [N, S] values never reach
        memory).  With ``posterior_variance_paths`` this is the variance explained by the data."""
        X = self._x(X)
        N = X.shape[0]
        with torch.cuda.device(self.device):
            out = torch.empty(N, dtype=F32, device=self.device)
            nbytes = C.c_size_t()
            _lib.check(self.lib.psbax_sumsq_scratch_bytes(C.byref(self._struct), N, C.byref(nbytes)),
                       "psbax_sumsq_scratch_bytes")
            scratch = torch.empty(max(nbytes.value, 4), dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.psbax_eval_sumsq(C.byref(self._struct), self._pack.data_ptr(), X.data_ptr(), N,
                                                 out.data_ptr(), scratch.data_ptr(), nbytes.value,
                                                 self._stream()), "psbax_eval_sumsq")
        return out

    def sumsq_groups(self, X, group: int) -> torch.Tensor:
        """``sum_{s in group g} f_s(X[i])**2``: [ceil(S / group), N] float32 on the GPU, one launch
        (``psbax_eval_sumsq_groups``; ``group`` divides 64)."""
        X = self._x(X)
        N = X.shape[0]
        with torch.cuda.device(self.device):
            out = torch.empty((self.S + group - 1) // group, N, dtype=F32, device=self.device)
            _lib.check(self.lib.psbax_eval_sumsq_groups(C.byref(self._struct), self._pack.data_ptr(), X.data_ptr(), N,
                                                        group, out.data_ptr(), self._stream()),
                       "psbax_eval_sumsq_groups")
        return out

    def levelset(self, X, threshold: float, want_mask: bool = True, row_offset: int = 0,
                 want_pooled: bool = False, mesh=None, N: Optional[int] = None) -> LevelSetResult:
        """Fused evaluation + ``fx > threshold`` for all S paths.

        ``want_pooled``: also return the OR of the S masks (``LevelSetResult.pooled``), formed inside the
        kernel — the pool ``gen_posterior_sampling_batch`` needs (``posterior_sampling.py:64-78``).
        ``mesh``: instead of ``X``, a list of d axis tensors; the candidates are rows
        [row_offset, row_offset + N) of ``reshape_mesh(torch.meshgrid(*axes, indexing='ij'))``
        (``src/utils.py:264-286``), generated on the device (``N`` defaults to the rest of the mesh)."""
        ms = None
        if mesh is not None:
            if X is not None:
                raise PsbaxError("pass either X or mesh")
            if self.input_transform is not None:
                raise PsbaxError("mesh candidates cannot go through an input transform")
            axes = [torch.as_tensor(ax).to(self.device, F32).contiguous().reshape(-1) for ax in mesh]
            if len(axes) != self.d or not 1 <= len(axes) <= 4:
                raise PsbaxError(f"mesh needs d={self.d} axes (at most 4), got {len(axes)}")
            total = math.prod(int(ax.numel()) for ax in axes)
            N = total - row_offset if N is None else int(N)
            ms = MeshStruct(d=len(axes), reserved=0)
            for i, ax in enumerate(axes):
                ms.len[i] = ax.numel()
                ms.axis[i] = ax.data_ptr()
            xptr = None
        else:
            X = self._x(X)
            N = X.shape[0]
            xptr = X.data_ptr()
        with torch.cuda.device(self.device):
            words = (N + 31) // 32
            mask = torch.empty(self.S, words, dtype=torch.int32, device=self.device) if want_mask else None
            pooled = torch.empty(words, dtype=torch.int32, device=self.device) if want_pooled else None
            count = torch.empty(self.S, dtype=torch.int64, device=self.device)
            maxval = torch.empty(self.S, dtype=F32, device=self.device)
            argmax = torch.empty(self.S, dtype=torch.int64, device=self.device)
            ws = self._workspace(N, 0)
            _lib.check(self.lib.psbax_eval_levelset_ex(
                C.byref(self._struct), self._pack.data_ptr(), xptr, C.byref(ms) if ms is not None else None, N,
                row_offset, float(threshold), _ptr(mask), words, _ptr(pooled), count.data_ptr(),
                maxval.data_ptr(), argmax.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()),
                "psbax_eval_levelset_ex")
        return LevelSetResult(mask, count, maxval, argmax, N, pooled)

    def levelset_streamed(self, X_host: torch.Tensor, threshold: float, chunk_rows: int = 1 << 21,
                          row_offset: int = 0):
        """Level set of a HOST-resident candidate set, pipelined: the rows are cut into chunks and
        chunk i+1 is uploaded (copy stream) while chunk i is evaluated (compute stream) and the
        mask of chunk i-1 is downloaded (second copy stream).  ``X_host`` should be pinned float32.

        Returns (mask_chunks, bounds, count, maxval, argmax): ``mask_chunks[c]`` is a pinned int32
        [S, ceil(rows_c / 32)] tensor for rows [bounds[c], bounds[c+1]); count / maxval / argmax
        are host tensors over the whole set.  The staging buffers (device double buffer, pinned
        result buffers) are cached per (N, chunk_rows) and reused: the returned tensors stay valid
        until the next call with the same shape."""
        X_host = torch.as_tensor(X_host)
        if X_host.device.type != "cpu" or X_host.dtype != F32 or X_host.dim() != 2 or X_host.shape[1] != self.d:
            raise PsbaxError("levelset_streamed expects a CPU float32 [N, d] tensor")
        N = X_host.shape[0]
        chunk_rows = max(4096, (chunk_rows // 2048) * 2048)
        bounds = _stream_bounds(N, chunk_rows)
        nchunk = len(bounds) - 1
        dev = self.device
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            if getattr(self, "_streams", None) is None:
                self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            s_in, s_out = self._streams
            cache = getattr(self, "_stream_cache", None)
            if cache is None or cache[0] != (N, chunk_rows, nchunk):
                xbuf = [torch.empty(chunk_rows, self.d, dtype=F32, device=dev) for _ in range(2)]
                masks_h = [torch.empty(self.S, (bounds[c + 1] - bounds[c] + 31) // 32, dtype=torch.int32).pin_memory()
                           for c in range(nchunk)]
                small = (torch.empty(self.S, dtype=torch.int64).pin_memory(),
                         torch.empty(self.S, dtype=F32).pin_memory(),
                         torch.empty(self.S, dtype=torch.int64).pin_memory())
                self._stream_cache = ((N, chunk_rows, nchunk), xbuf, masks_h, small)
            _, xbuf, masks_h, (cnt_h, bv_h, bi_h) = self._stream_cache
            ev_in = [torch.cuda.Event() for _ in range(nchunk)]
            ev_done = [torch.cuda.Event() for _ in range(nchunk)]
            ev_free = [None, None]  # compute finished with xbuf[b]
            s_in.wait_stream(main)
            s_out.wait_stream(main)
            results = []
            for c in range(nchunk):
                lo, hi = bounds[c], bounds[c + 1]
                b = c & 1
                with torch.cuda.stream(s_in):
                    if ev_free[b] is not None:
                        s_in.wait_event(ev_free[b])
                    xbuf[b][: hi - lo].copy_(X_host[lo:hi], non_blocking=True)
                    ev_in[c].record(s_in)
                main.wait_event(ev_in[c])
                r = self.levelset(xbuf[b][: hi - lo], threshold, True, row_offset + lo)
                ev_free[b] = torch.cuda.Event()
                ev_free[b].record(main)
                ev_done[c].record(main)
                results.append(r)  # keeps the device mask alive until its download is queued
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done[c])
                    masks_h[c].copy_(r.mask, non_blocking=True)
                    r.mask.record_stream(s_out)
            # combine the per-chunk results once (no small kernels between the chunk launches);
            # ties go to the earlier chunk, i.e. to the smaller row index
            tot = torch.stack([r.count for r in results]).sum(0)
            best_v, which = torch.stack([r.maxval for r in results]).max(0)
            best_i = torch.stack([r.argmax for r in results]).gather(0, which[None])[0]
            cnt_h.copy_(tot, non_blocking=True)
            bv_h.copy_(best_v, non_blocking=True)
            bi_h.copy_(best_i, non_blocking=True)
            main.wait_stream(s_out)
            main.synchronize()
        return masks_h, bounds, cnt_h, bv_h, bi_h

    def argmax(self, X, row_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """Per-sample max and arg max over the candidates (ES-population scoring)."""
        r = self.levelset(X, float("inf"), want_mask=False, row_offset=row_offset)
        return r.maxval, r.argmax

    def topk(self, X, k: int, row_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """Fused evaluation + per-sample top-k: (values [S, k], indices [S, k])."""
        X = self._x(X)
        N = X.shape[0]
        with torch.cuda.device(self.device):
            val = torch.empty(self.S, k, dtype=F32, device=self.device)
            idx = torch.empty(self.S, k, dtype=torch.int64, device=self.device)
            ws = self._workspace(N, k)
            _lib.check(self.lib.psbax_eval_topk(
                C.byref(self._struct), self._pack.data_ptr(), X.data_ptr(), N, row_offset, k,
                val.data_ptr(), idx.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()),
                "psbax_eval_topk")
        return val, idx

    def grad(self, X, sample_idx=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(f_{s_i}(X[i]), d f_{s_i} / dx (X[i])) for (point, sample) pairs."""
        X = self._x(X)
        N = X.shape[0]
        with torch.cuda.device(self.device):
            si = None
            if sample_idx is not None:
                si = torch.as_tensor(sample_idx).to(self.device, torch.int32).contiguous().reshape(-1)
                if si.numel() != N:
                    raise PsbaxError("sample_idx must have one entry per row of X")
            val = torch.empty(N, dtype=F32, device=self.device)
            g = torch.empty(N, self.d, dtype=F32, device=self.device)
            _lib.check(self.lib.psbax_eval_grad(C.byref(self._struct), self._pack.data_ptr(),
                                                X.data_ptr(), N, _ptr(si), val.data_ptr(),
                                                g.data_ptr(), self._stream()), "psbax_eval_grad")
        return val, g

    def ascend(self, X0, sample_idx=None, maxiter: int = 100, pgtol: Optional[float] = None,
               ftol: float = 2e-7) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """P independent box-constrained ascents on [0, 1]^d in one launch (``psbax_ascend``): ascent i
        maximises path ``sample_idx[i]`` from ``X0[i]``.  Returns (x [P, d], f(x) [P], iterations [P])."""
        if self.input_transform is not None:
            raise PsbaxError("ascent through an input transform is not implemented")
        X0 = self._x(X0)
        P = X0.shape[0]
        with torch.cuda.device(self.device):
            si = None
            if sample_idx is not None:
                si = torch.as_tensor(sample_idx).to(self.device, torch.int32).contiguous().reshape(-1)
                if si.numel() != P:
                    raise PsbaxError("sample_idx must have one entry per row of X0")
            x = torch.empty(P, self.d, dtype=F32, device=self.device)
            f = torch.empty(P, dtype=F32, device=self.device)
            it = torch.empty(P, dtype=torch.int32, device=self.device)
            pg = 1e-5 * abs(self.y_std) if pgtol is None else float(pgtol)
            _lib.check(self.lib.psbax_ascend(C.byref(self._struct), self._pack.data_ptr(), X0.data_ptr(), P, _ptr(si),
                                             int(maxiter), pg, float(ftol), x.data_ptr(), f.data_ptr(), it.data_ptr(),
                                             self._stream()), "psbax_ascend")
        return x, f, it

    def subset_select(self, X, etas, k: int, nonneg: bool = False, noise: str = "additive") -> torch.Tensor:
        """DiscoBAX greedy subset per sample (``src/algorithms/discobax.py:37-56``): [S, k] indices.
        ``etas``: [B, N] (one noise matrix for all paths) or [S, B, N] (one per path); ``noise``:
        "additive" (values = fx + etas) or "multiplicative" (values = fx * etas, etas the 0 / 1 masks of
        ``src/problems.py:150-160``)."""
        fx = self.values(X)
        N = fx.shape[0]
        etas = torch.as_tensor(etas).to(self.device, F32).contiguous()
        if etas.dim() not in (2, 3) or etas.shape[-1] != N or (etas.dim() == 3 and etas.shape[0] != self.S):
            raise PsbaxError("etas must be [B, N] or [S, B, N]")
        if noise not in ("additive", "multiplicative"):
            raise PsbaxError(f"noise model {noise!r}")
        B = etas.shape[-2]
        with torch.cuda.device(self.device):
            out = torch.empty(self.S, k, dtype=torch.int64, device=self.device)
            _lib.check(self.lib.psbax_subset_select_ex(fx.data_ptr(), self.S, etas.data_ptr(),
                                                       B * N if etas.dim() == 3 else 0, N, B, self.S, k,
                                                       int(bool(nonneg)), int(noise == "multiplicative"),
                                                       out.data_ptr(), self._stream()),
                       "psbax_subset_select_ex")
        return out

    def path(self, s: int) -> "SamplePath":
        return SamplePath(self, s)

    def __len__(self) -> int:
        return self.S

    def __iter__(self):
        return (self.path(s) for s in range(self.S))


class MultiOutputSamplePathBatch:
    """S sample paths of an m-output ``ModelListGP`` (``src/utils.py:231-248``): one independent
    ``SamplePathBatch`` per output model (its own basis, lengthscales and update weights), evaluated
    over the same candidates and concatenated on the last dimension, as the reference's
    ``aux_func`` does with the m ``get_gp_samples`` models."""

    def __init__(self, batches):
        self.batches = list(batches)
        if len({b.S for b in self.batches}) != 1 or len({b.d for b in self.batches}) != 1:
            raise PsbaxError("every output needs the same number of paths over the same input space")
        self.S, self.d, self.m = self.batches[0].S, self.batches[0].d, len(self.batches)
        self.device = self.batches[0].device

    def values(self, X) -> torch.Tensor:
        """[N, S, m] float32 on the GPU: m fused launches, one per output model."""
        return torch.stack([b.values(X) for b in self.batches], dim=-1)

    def output(self, j: int) -> SamplePathBatch:
        return self.batches[j]

    def path(self, s: int) -> "MultiOutputSamplePath":
        return MultiOutputSamplePath(self, s)

    def __len__(self) -> int:
        return self.S

    def __iter__(self):
        return (self.path(s) for s in range(self.S))


class MultiOutputSamplePath:
    """One m-output sample with the calling convention of the ``GenericDeterministicModel`` the
    reference returns for a ``ModelListGP`` (``src/utils.py:242-248``): X [b, 1, d] -> [b, 1, m]
    (X [b, d] -> [b, m]); ``.posterior(X).mean`` gives the same tensor."""

    def __init__(self, batch: MultiOutputSamplePathBatch, s: int):
        self.batch, self.s = batch, s
        self.paths = [b.path(s) for b in batch.batches]

    def __call__(self, X) -> torch.Tensor:
        X = torch.as_tensor(X)
        out = torch.stack([p(X) for p in self.paths], dim=-1)  # [b, m]
        return out.unsqueeze(1) if X.dim() == 3 else out

    def posterior(self, X):
        from types import SimpleNamespace

        return SimpleNamespace(mean=self(X))


class _PathValue(torch.autograd.Function):
    """f_s(X) with the CUDA gradient kernel as backward (LBFGSB needs d f / d x,
    ``src/algorithms/lbfgsb.py:29-39``)."""

    @staticmethod
    def forward(ctx, X, path):
        val, g = path.batch.grad(X.detach().reshape(-1, path.batch.d),
                                 torch.full((X.numel() // path.batch.d,), path.s, dtype=torch.int32))
        ctx.save_for_backward(g.to(X.device, X.dtype).reshape(X.shape))
        return val.to(X.device, X.dtype)

    @staticmethod
    def backward(ctx, gout):
        (g,) = ctx.saved_tensors
        return g * gout.reshape(-1, *([1] * (g.dim() - 1))), None


class SamplePath:
    """One sample path with the calling convention of the object
    ``get_function_samples`` returns for a ``SingleTaskGP``
    (``PosteriorMean(sample)``, ``src/utils.py:229``): X [b, 1, d] -> [b];
    X [b, d] -> [b].  Output follows X's dtype/device."""

    def __init__(self, batch: SamplePathBatch, s: int):
        if not 0 <= s < batch.S:
            raise IndexError(s)
        self.batch, self.s = batch, s
        self._single = None

    def _one(self) -> SamplePathBatch:
        # a one-sample view so that repeated legacy calls do not evaluate all S paths
        if self._single is None:
            b, s = self.batch, self.s
            if b.S == 1:
                self._single = b
            else:
                g = 0 if b.G == 1 else s
                self._single = SamplePathBatch(
                    b.omega[g:g + 1], b.bias[g:g + 1], b.W[s:s + 1], b.kernel, b.Xtrain, b.inv_ls,
                    None if b.V is None else b.V[s:s + 1], b.mean_const, b.y_mean, b.y_std,
                    b.device, "simt", b.input_transform)
        return self._single

    def __call__(self, X) -> torch.Tensor:
        X = torch.as_tensor(X)
        d = X.shape[-1] if self.batch.input_transform is not None else self.batch.d
        if X.dim() == 3 and X.shape[1] != 1:
            raise PsbaxError("expected X of shape [b, 1, d] (q = 1)")
        if X.requires_grad:
            if self.batch.input_transform is not None:
                raise PsbaxError("gradients through an input transform are not implemented")
            flat = X.reshape(-1, d)
            return _PathValue.apply(flat, self)
        out = self._one().values(X.reshape(-1, d))[:, 0]
        dtype = X.dtype if X.dtype.is_floating_point else F64
        return out.to(device=X.device, dtype=dtype)


# ------------------------------------------------------------------------------------
# Parameter draws (torch; the n x n / L x L factorisations stay in cuSOLVER via torch)
# ------------------------------------------------------------------------------------
def _basis(d, L, kernel, G, gen, dev):
    """botorch ``RandomFourierFeatures.__init__``: randn(d, L),
    Matern columns scaled by rsqrt(Gamma(nu, nu)), bias = 2 pi rand(L); G independent bases."""
    w = torch.randn(G, d, L, dtype=F64, device=dev, generator=gen)
    if kernel != "rbf":
        nu = {"matern12": 0.5, "matern32": 1.5, "matern52": 2.5}[kernel]
        g = torch._standard_gamma(torch.full((G, 1, L), nu, dtype=F64, device=dev), generator=gen) / nu
        w = torch.rsqrt(g) * w
    b = 2.0 * math.pi * torch.rand(G, L, dtype=F64, device=dev, generator=gen)
    return w, b


def _model_state(model, dev):
    X = model.train_inputs[0].to(dev, F64)
    y = model.train_targets.to(dev, F64)
    ls = model.lengthscale_vector().to(dev, F64)
    alpha = float(model.covar_module.outputscale)
    noise = float(model.likelihood.noise.mean())
    mu = float(model.mean_module.constant)
    y_mean, y_std = model.y_mean_std()
    return X, y, ls, alpha, noise, mu, y_mean, y_std


def matheron_params(model, S: int, L: int = 1000, generator: Optional[torch.Generator] = None,
                    device=None) -> dict:
    """Parameters of S pathwise (Matheron) posterior samples, the north-star form: shared
    RFF basis, prior weights w_s ~ N(0, I), exact-kernel update
    v_s = (K + sigma^2 I)^-1 (y - mu - Phi w_s - eps_s).  One n x n Cholesky for all S.
    Pure torch (float64) on ``device``; returns the keyword arguments of ``SamplePathBatch``."""
    dev = torch.device(device) if device is not None else default_device()
    X, y, ls, alpha, noise, mu, y_mean, y_std = _model_state(model, dev)
    n, d = X.shape
    kernel = model.kernel
    wts, b = _basis(d, L, kernel, 1, generator, dev)
    omega = (wts[0] / ls[:, None]).T.contiguous()  # [L, d]
    scale = math.sqrt(2.0 * alpha / L)
    Phi = scale * torch.cos(X @ omega.T + b[0])  # [n, L]
    w = torch.randn(S, L, dtype=F64, device=dev, generator=generator)
    eps = math.sqrt(noise) * torch.randn(S, n, dtype=F64, device=dev, generator=generator)
    Lk = model.train_cholesky().to(dev)
    resid = y[None, :] - mu - w @ Phi.T - eps
    v = torch.cholesky_solve(resid.T, Lk).T  # [S, n]
    return dict(omega=omega[None], bias=b, W=scale * w, kernel=kernel, Xtrain=X / ls,
                inv_ls=1.0 / ls, V=alpha * v, mean_const=mu, y_mean=y_mean, y_std=y_std)


def reference_params(model, S: int, L: int = 1000, generator: Optional[torch.Generator] = None,
                     device=None) -> dict:
    """Parameters of S samples drawn with the reference's scheme (``src/utils.py:222-229`` ->
    botorch ``get_gp_samples``): a fresh basis per sample and w_s ~ N(m, sigma^2 A^-1),
    A = Phi^T Phi + sigma^2 I; no exact-kernel term; the GP constant mean is not used.
    Batched over the samples."""
    dev = torch.device(device) if device is not None else default_device()
    X, y, ls, alpha, noise, _mu, y_mean, y_std = _model_state(model, dev)
    n, d = X.shape
    kernel = model.kernel
    wts, b = _basis(d, L, kernel, S, generator, dev)
    omega = (wts / ls[None, :, None]).transpose(1, 2).contiguous()  # [S, L, d]
    scale = math.sqrt(2.0 * alpha / L)
    Ws = []
    eye = torch.eye(L, dtype=F64, device=dev)
    for s0 in range(0, S, 16):  # chunks bound the S x L x L workspace
        om = omega[s0:s0 + 16]
        Phi = scale * torch.cos(torch.einsum("nd,sld->snl", X, om) + b[s0:s0 + 16, None, :])
        A = Phi.transpose(1, 2) @ Phi + noise * eye
        LA = torch.linalg.cholesky(A)
        rhs = Phi.transpose(1, 2) @ y[None, :, None].expand(om.shape[0], n, 1)
        m = torch.cholesky_solve(rhs, LA)[..., 0]
        # w = m + sqrt(noise) * LA^-T z has covariance noise * A^-1
        z = torch.randn(om.shape[0], L, 1, dtype=F64, device=dev, generator=generator)
        u = torch.linalg.solve_triangular(LA.transpose(1, 2), z, upper=True)[..., 0]
        Ws.append(m + math.sqrt(noise) * u)
    return dict(omega=omega, bias=b, W=scale * torch.cat(Ws), kernel=kernel, Xtrain=None,
                inv_ls=None, V=None, mean_const=0.0, y_mean=y_mean, y_std=y_std)


def posterior_mean_params(model, device=None) -> dict:
    """``PosteriorMean(model)`` as a one-sample path with L = 0:
    mu(x) = mean_const + k(x, X) (K + sigma^2 I)^-1 (y - mean_const)
    (what ``src/performance_metrics.py:117-122`` feeds the algorithms)."""
    dev = torch.device(device) if device is not None else default_device()
    X, y, ls, alpha, noise, mu, y_mean, y_std = _model_state(model, dev)
    d = X.shape[1]
    Lk = model.train_cholesky().to(dev)
    a = torch.cholesky_solve((y - mu)[:, None], Lk)[:, 0]
    return dict(omega=torch.zeros(1, 0, d), bias=torch.zeros(1, 0), W=torch.zeros(1, 0),
                kernel=model.kernel, Xtrain=X / ls, inv_ls=1.0 / ls, V=(alpha * a)[None],
                mean_const=mu, y_mean=y_mean, y_std=y_std)


def posterior_variance_params(model, pending: Optional[torch.Tensor] = None, jitter: float = 1e-8,
                              device=None, pending_noise=None) -> dict:
    """Pseudo-paths whose row-wise squared norm is the variance explained by the data
    (SURVEY.md section 8, row f1):  with K + sigma^2 I = L L^T,

        sum_s f_s(x)^2 = | L^-1 k(X, x) |^2 = k(x, X) (K + sigma^2 I)^-1 k(X, x),

    f_s(x) = sum_i V[s, i] k(x, x_i) / alpha,  V = alpha L^-1  (S = n rows, L = 0 RFF features), so
    var_post(x) = alpha - sumsq(x) in standardized units.  ``pending`` rows [m, d] are appended as
    noise-free observations (plus ``jitter``): the variance is then conditional on them, which is
    what the greedy q > 1 loop of ``optimize_acqf_discrete`` maximises
    (logdet Cov([x, P]) = logdet Cov(P) + log var(x | P)).  ``pending_noise`` (scalar or [m], in
    standardized units) replaces the jitter: rows observed WITH noise, e.g. the execution path a
    fantasy model of INFO-BAX is conditioned on (``bax_acquisition.py:177-203``)."""
    dev = torch.device(device) if device is not None else default_device()
    X, y, ls, alpha, noise, mu, y_mean, y_std = _model_state(model, dev)
    n, d = X.shape
    diag = torch.full((n,), float(noise), dtype=F64, device=dev)
    if pending is not None and pending.numel() > 0:
        P = pending.reshape(-1, d).to(dev, F64)
        X = torch.cat([X, P])
        if pending_noise is None:
            extra = torch.full((P.shape[0],), float(jitter) * float(alpha), dtype=F64, device=dev)
        else:
            extra = torch.as_tensor(pending_noise, dtype=F64, device=dev).reshape(-1).expand(P.shape[0]).clone()
            extra = extra.clamp_min(float(jitter) * float(alpha))
        diag = torch.cat([diag, extra])
    from .models.gp import kernel_profile

    a = X / ls
    K = alpha * kernel_profile(((a[:, None, :] - a[None, :, :]) ** 2).sum(-1), model.kernel) + torch.diag(diag)
    Lk = torch.linalg.cholesky(K)
    Linv = torch.linalg.solve_triangular(Lk, torch.eye(X.shape[0], dtype=F64, device=dev), upper=False)
    S = X.shape[0]
    return dict(omega=torch.zeros(1, 0, d), bias=torch.zeros(1, 0), W=torch.zeros(S, 0), kernel=model.kernel,
                Xtrain=X / ls, inv_ls=1.0 / ls, V=alpha * Linv, mean_const=0.0, y_mean=0.0, y_std=1.0)


def fantasy_variance_params(model, path_xs, pending: Optional[torch.Tensor] = None, jitter: float = 1e-8,
                            group: int = 16, device=None):
    """Pseudo-paths for the posterior variance of the current model AND of one fantasy model per
    execution path, over one candidate stream (SURVEY.md section 8, row f3;
    ``src/acquisition_functions/bax_acquisition.py:87-122, 177-203``).

    Base system: Z = [X; pending], K0 = k(Z, Z) + diag(sigma^2 .. | jitter ..) = L0 L0^T.  The fantasy of
    path s (rows P_s, observed with the likelihood noise) has the block Cholesky factor
    [[L0, 0], [B_s, L_s]],  B_s = k(P_s, Z) L0^-T,  L_s L_s^T = k(P_s, P_s) + sigma^2 I - B_s B_s^T, whose
    inverse only ADDS the rows [-L_s^-1 B_s L0^-1, L_s^-1] to L0^-1.  So with the training set
    [Z; P_1; ..; P_S] the pseudo-paths are: the n0 base rows (group-padded), then per path its k_s rows
    (group-padded), and

        var_0(x) = alpha - sum_{base groups} sumsq_g(x),     var_s(x) = var_0(x) - sum_{groups of s} sumsq_g(x).

    Returns (SamplePathBatch kwargs, number of base groups, [groups per path])."""
    dev = torch.device(device) if device is not None else default_device()
    X, y, ls, alpha, noise, mu, y_mean, y_std = _model_state(model, dev)
    from .models.gp import kernel_profile

    def kmat(A, B):
        a, b = A / ls, B / ls
        return alpha * kernel_profile(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1), model.kernel)

    d = X.shape[1]
    Z, diag = X, torch.full((X.shape[0],), float(noise), dtype=F64, device=dev)
    if pending is not None and pending.numel() > 0:
        P = pending.reshape(-1, d).to(dev, F64)
        Z = torch.cat([Z, P])
        diag = torch.cat([diag, torch.full((P.shape[0],), float(jitter) * float(alpha), dtype=F64, device=dev)])
    n0 = Z.shape[0]
    L0 = torch.linalg.cholesky(kmat(Z, Z) + torch.diag(diag))
    L0inv = torch.linalg.solve_triangular(L0, torch.eye(n0, dtype=F64, device=dev), upper=False)
    paths = [torch.as_tensor(p).reshape(-1, d).to(dev, F64) for p in path_xs]
    pad = lambda k: ((k + group - 1) // group) * group  # noqa: E731
    g0 = pad(n0) // group
    gs = [pad(p.shape[0]) // group for p in paths]
    n_total = n0 + sum(p.shape[0] for p in paths)
    V = torch.zeros((g0 + sum(gs)) * group, n_total, dtype=F64, device=dev)
    V[:n0, :n0] = L0inv
    row, col = g0 * group, n0
    for p, g in zip(paths, gs):
        k = p.shape[0]
        B = kmat(p, Z) @ L0inv.T
        Cs = kmat(p, p) + float(noise) * torch.eye(k, dtype=F64, device=dev) - B @ B.T
        Ls = torch.linalg.cholesky(0.5 * (Cs + Cs.T))
        Lsinv = torch.linalg.solve_triangular(Ls, torch.eye(k, dtype=F64, device=dev), upper=False)
        V[row:row + k, :n0] = -Lsinv @ B @ L0inv
        V[row:row + k, col:col + k] = Lsinv
        row, col = row + g * group, col + k
    Xall = torch.cat([Z] + paths)
    S = V.shape[0]
    params = dict(omega=torch.zeros(1, 0, d), bias=torch.zeros(1, 0), W=torch.zeros(S, 0), kernel=model.kernel,
                  Xtrain=Xall / ls, inv_ls=1.0 / ls, V=alpha * V, mean_const=0.0, y_mean=0.0, y_std=1.0)
    return params, g0, gs


def posterior_variance(model, Xq, pending: Optional[torch.Tensor] = None, device=None,
                       engine: Optional[str] = None, pending_noise=None) -> torch.Tensor:
    """Exact-GP posterior variance (original units, latent f) at every row of ``Xq`` [N, d] in one
    fused launch: y_std^2 (alpha - sumsq(Xq)).  [N] float32 on the GPU.

    alpha - |L^-1 k|^2 cancels where the data explain most of the prior variance, so the default
    engine is the 3xTF32 split (measured max abs error 7e-6 alpha against 1.2e-4 alpha for BF16x3
    at n = 64, noise 1e-4; ``scripts/var_probe.py``) wherever that engine supports the shape."""
    dev = torch.device(device) if device is not None else default_device()
    if getattr(model, "dkl", False):
        return _dkgp_posterior_variance(model, Xq, pending, dev, engine)
    params = posterior_variance_params(model, pending, device=dev, pending_noise=pending_noise)
    if engine is None:
        try:
            batch = SamplePathBatch(**params, device=dev, engine="tensor")
        except _lib.PsbaxError as ex:  # operands too large for the TF32 images: the FFMA engine (full fp32
            if "engine needs" not in str(ex):  # products), never the BF16x3 split, for a difference that cancels
                raise
            batch = SamplePathBatch(**params, device=dev, engine="simt")
    else:
        batch = SamplePathBatch(**params, device=dev, engine=engine)
    alpha = float(model.covar_module.outputscale)
    _, y_std = model.y_mean_std()
    return (alpha - batch.sumsq(Xq)) * (float(y_std) ** 2)


def _embedder(model, dev):
    """The model's embedding as a GPU op that leaves the caller's network where it is."""
    net = getattr(model, "feature_extractor", None)
    if net is None:
        return lambda X: torch.as_tensor(X).to(dev, F32)
    return DeviceFFNN(net, dev)


def _dkgp_posterior_variance(model, Xq, pending, dev, engine) -> torch.Tensor:
    """Deep-kernel GP (linear kernel on the embedding): var(x) = e(x)^T Sn e(x) = |Ls^T e(x)|^2 with
    Sn = Ls Ls^T the weight-space posterior covariance (conditioned on the pending rows): h linear-
    kernel pseudo-paths V = Ls^T over the embedded candidates — ``psbax_eval_sumsq`` gives the
    variance directly."""
    _, Sn = model.weight_posterior(pending)
    Ls = torch.linalg.cholesky(Sn.to(dev))
    h = Ls.shape[0]
    batch = SamplePathBatch(omega=torch.zeros(1, 0, h), bias=torch.zeros(1, 0), W=torch.zeros(h, 0), kernel="linear",
                            Xtrain=torch.eye(h, dtype=F64), inv_ls=torch.ones(h, dtype=F64), V=Ls.T.contiguous(),
                            mean_const=0.0, y_mean=0.0, y_std=1.0, device=dev, engine=engine or "auto",
                            input_transform=_embedder(model, dev))
    return batch.sumsq(Xq)


def dkgp_linear_params(model, S: int, generator: Optional[torch.Generator] = None, device=None) -> dict:
    """DKGP with a linear kernel on the network embedding (``src/utils.py:153-206``): exact
    weight-space draw  Sn^-1 = I / v + E^T E / sigma^2,  mn = Sn E^T (y - mu) / sigma^2,
    w_s ~ N(mn, Sn),  f_s(x) = emb(x) . w_s + mu.  Expressed for the evaluator as L = 0 and h
    linear-kernel rows Xtrain = I_h, V = W (kernel "linear"); the embedding is the batch's input
    transform."""
    dev = torch.device(device) if device is not None else default_device()
    with torch.no_grad():
        E = model.embedding(model.train_inputs[0]).to(dev, F64)  # [n, h] (float64, wherever the network lives)
        y = model.train_targets.to(dev, F64)
        h = E.shape[1]
        v = float(model.covar_module.base_kernel.variance)
        sigma2 = float(model.likelihood.noise.reshape(-1)[0])
        mu = float(model.mean_module.constant)
        Sn_inv = torch.eye(h, dtype=F64, device=dev) / v + E.T @ E / sigma2
        mn = torch.linalg.solve(Sn_inv, E.T @ (y - mu)) / sigma2
        Lp = torch.linalg.cholesky(Sn_inv)  # Sn = (Lp Lp^T)^-1  ->  w = mn + Lp^-T z
        z = torch.randn(h, S, dtype=F64, device=dev, generator=generator)
        W = (mn[:, None] + torch.linalg.solve_triangular(Lp.T, z, upper=True)).T  # [S, h]
    return dict(omega=torch.zeros(1, 0, h), bias=torch.zeros(1, 0), W=torch.zeros(S, 0), kernel="linear",
                Xtrain=torch.eye(h, dtype=F64), inv_ls=torch.ones(h, dtype=F64), V=W, mean_const=mu,
                y_mean=0.0, y_std=1.0)


def draw_dkgp_linear_paths(model, S: int, generator=None, device=None, engine: str = "auto") -> SamplePathBatch:
    """The embedding of all N candidates runs in ``psbax_mlp_forward`` (float32, a device copy of the
    network's weights: the caller's model stays on its device), fused evaluation + reduction after it."""
    dev = torch.device(device) if device is not None else default_device()
    return SamplePathBatch(**dkgp_linear_params(model, S, generator, dev), device=dev, engine=engine,
                           input_transform=_embedder(model, dev))


def draw_dkgp_rff_paths(model, S: int, L: int = 1000, generator=None, device=None, engine: str = "auto") -> SamplePathBatch:
    """DKGP with an RBF / Matern kernel on the embedding (``src/utils.py:207-221``): the reference copies
    the model, replaces its training inputs by their embedding and draws an RFF sample of that GP, then
    evaluates it on ``embedding(X)``.  Same here with the Matheron form on the embedded training set;
    the embedding is the batch's input transform (``psbax_mlp_forward``)."""
    from .models.gp import SingleTaskGP

    dev = torch.device(device) if device is not None else default_device()
    with torch.no_grad():
        E = model.embedding(model.train_inputs[0]).to("cpu", F64)
    bk = model.covar_module.base_kernel
    aux = SingleTaskGP(E, model.train_targets.to("cpu", F64), kernel=bk.kind, lengthscale=bk.lengthscale.to("cpu"),
                       outputscale=float(model.covar_module.outputscale), noise=float(model.likelihood.noise.mean()),
                       mean_const=float(model.mean_module.constant), standardize=False)  # targets standardized by fit_model
    batch = SamplePathBatch(**matheron_params(aux, S, L, generator, dev), device=dev, engine=engine,
                            input_transform=_embedder(model, dev))
    return batch


def draw_matheron_paths(model, S: int, L: int = 1000, generator=None, device=None,
                        engine: str = "auto") -> SamplePathBatch:
    dev = torch.device(device) if device is not None else default_device()
    return SamplePathBatch(**matheron_params(model, S, L, generator, dev), device=dev, engine=engine)


def draw_reference_paths(model, S: int, L: int = 1000, generator=None, device=None) -> SamplePathBatch:
    dev = torch.device(device) if device is not None else default_device()
    return SamplePathBatch(**reference_params(model, S, L, generator, dev), device=dev, engine="simt")


def posterior_mean_path(model, device=None) -> SamplePathBatch:
    dev = torch.device(device) if device is not None else default_device()
    return SamplePathBatch(**posterior_mean_params(model, dev), device=dev, engine="simt")

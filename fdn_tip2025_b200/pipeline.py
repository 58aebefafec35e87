"""Device-resident version of the reference's inference loop (inference_fdn_lolblur.py:42-75, inference_fdn_lolv1.py:39-68).

The scripts do, per image on the host: cv2.imread -> float32/255 -> img2tensor(bgr2rgb) -> .to(device) -> reflect-pad to a multiple
of 32 -> LPNet -> FDN -> crop -> tensor2img (clamp, *255, round, uint8, RGB->BGR) -> imwrite.  Here only the uint8 frame crosses
PCIe in either direction (4x fewer bytes than fp32) and every step in between is a kernel of libfdn_b200.so:

    pipe = InferencePipeline(net, net_ipred, variant="lolblur")        # or "lolv1"
    out_bgr = pipe(cv2.imread(path))                                   # uint8 [h,w,3] in, uint8 [h,w,3] out

``variant`` selects what the script passes as ``ratio_i``: the LPNet output itself (lolblur, inference_fdn_lolblur.py:65,71) or
mean(gray(img)) / LPNet(img) (lolv1, inference_fdn_lolv1.py:57-64).
"""
import threading

import numpy as np
import torch

from . import ops

# Graph capture is serialised across host threads: a device-wide synchronisation (the warm-up's, or the one-off upload of an FFT
# twiddle table) issued by one worker while another captures on the same device is an error (cudaErrorStreamCaptureUnsupported).
_CAPTURE_LOCK = threading.Lock()


def padded_size(h, w, multiple=32):
    return h + (multiple - h % multiple) % multiple, w + (multiple - w % multiple) % multiple


class InferencePipeline:
    def __init__(self, net, net_ipred, variant="lolblur", use_graphs=False):
        if variant not in ("lolblur", "lolv1"):
            raise ValueError("variant must be 'lolblur' or 'lolv1'")
        self.net, self.net_ipred, self.variant, self.use_graphs = net, net_ipred, variant, use_graphs
        self.device = next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("the pipeline runs on a CUDA device; there is no CPU path")

    @torch.no_grad()
    def run_device(self, frames_u8):
        """frames_u8: uint8 CUDA tensor [B,h,w,3] (BGR).  Returns (uint8 CUDA tensor [B,h,w,3] BGR, ratio_i [B,1])."""
        b, h, w, c = frames_u8.shape
        if c != 3:
            raise RuntimeError("expected [B,h,w,3] BGR frames")
        hp, wp = padded_size(h, w)
        x = torch.empty(b, 3, hp, wp, dtype=torch.float32, device=frames_u8.device)
        ops.pre_u8hwc(frames_u8.contiguous(), x)
        ratio = self.net_ipred(x)
        if self.variant == "lolv1":
            gray = torch.empty(b, dtype=torch.float32, device=x.device)
            ops.gray_mean(x, gray)
            ratio = gray.view(b, 1) / ratio
        restored = self.net(x, ratio_i=ratio)[0]
        out = torch.empty(b, h, w, 3, dtype=torch.uint8, device=x.device)
        ops.post_u8hwc(restored, out)
        return out, ratio

    # ---- CUDA-graph replay -------------------------------------------------------------------
    # A forward is ~750 kernel launches; for small frames (256x256, 400x600) the host cannot issue them as fast as the GPU
    # retires them.  Per frame shape the whole device-side sequence (pre -> LPNet -> FDN -> post) is captured once into a CUDA
    # graph with static uint8 input / output buffers and replayed afterwards: one launch per frame.
    def _graph_for(self, b, h, w):
        key = (b, h, w)
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        entry = self._graphs.get(key)
        if entry is None:
            with _CAPTURE_LOCK:
                entry = self._capture(key, b, h, w)
        return entry

    def _capture(self, key, b, h, w):
        entry = None
        if entry is None:
            frames = torch.zeros(b, h, w, 3, dtype=torch.uint8, device=self.device)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):               # warm-up outside capture: twiddle tables, packed weights, kernel attributes
                for _ in range(2):
                    self.run_device(frames)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            # an explicit capture stream on THIS device: torch.cuda.graph's default capture stream is created once per process, on
            # whichever device was current first, and kernels captured on another GPU's stream fault
            with torch.cuda.graph(graph, stream=side, capture_error_mode="relaxed"):    # kernel-attribute calls of the C ABI are not stream work
                out, ratio = self.run_device(frames)
            entry = (graph, frames, out, ratio)
            self._graphs[key] = entry
        return entry

    @torch.no_grad()
    def run_device_graphed(self, frames_u8):
        """Same contract as run_device, replaying a CUDA graph captured for this (B, h, w).  The returned tensors are the graph's
        static buffers: consume (or copy) them before the next call with the same shape."""
        b, h, w, _ = frames_u8.shape
        graph, frames, out, ratio = self._graph_for(b, h, w)
        frames.copy_(frames_u8, non_blocking=True)
        graph.replay()
        return out, ratio

    def __call__(self, img_bgr_u8):
        """numpy uint8 [h,w,3] (or [B,h,w,3]) BGR as cv2.imread returns it -> restored frame(s), same layout."""
        a = np.ascontiguousarray(img_bgr_u8)
        if a.dtype != np.uint8:
            raise RuntimeError("expected the uint8 array cv2.imread returns")
        single = a.ndim == 3
        t = torch.from_numpy(a[None] if single else a).pin_memory().to(self.device, non_blocking=True)
        out, _ = self.run_device_graphed(t) if self.use_graphs else self.run_device(t)
        res = out.cpu().numpy()
        return res[0] if single else res


# =====================================================================================================================================
# Folder / batch driver over several GPUs (replaces the batch-1 synchronous loop of
# inference_fdn_lolblur.py:42-75 and the `idx % world_size != rank` sharding of image_restoration_model.py:728-732)
# =====================================================================================================================================
import queue
import threading


def plan_batches(shapes, budget_pixels):
    """Group frame indices into micro-batches: frames of one batch share a shape (one kernel sequence / CUDA graph per shape) and a
    batch holds at most ``budget_pixels`` padded pixels (the activations of a forward are ~2.4 KB per padded pixel, so the budget is a
    memory bound).  ``shapes``: list of (h, w).  Returns a list of index lists; the first-seen order of frames is kept inside a shape."""
    by_shape = {}
    for i, (h, w) in enumerate(shapes):
        by_shape.setdefault((h, w), []).append(i)
    batches = []
    for (h, w), idx in by_shape.items():
        hp, wp = padded_size(h, w)
        per = max(1, int(budget_pixels // (hp * wp)))
        for s in range(0, len(idx), per):
            batches.append(idx[s:s + per])
    return batches


def shard_frames(n_frames, n_workers):
    """Frame i -> worker i mod G (image_restoration_model.py:731)."""
    return [list(range(r, n_frames, n_workers)) for r in range(n_workers)]


class _Worker(threading.Thread):
    """One host thread per GPU: owns the module replicas, one stream and pinned uint8 staging buffers on its device."""

    def __init__(self, index, device, make_nets, variant, budget_pixels, use_graphs):
        super().__init__(daemon=True, name="fdn-worker-%d" % index)
        self.index, self.device, self.variant = index, torch.device(device), variant
        self.make_nets, self.budget, self.use_graphs = make_nets, budget_pixels, use_graphs
        self.jobs, self.error = queue.Queue(), None
        self.ready = threading.Event()
        self.pinned = {}

    def _pinned(self, kind, shape):
        key = (kind,) + tuple(shape)
        buf = self.pinned.get(key)
        if buf is None:
            buf = torch.empty(shape, dtype=torch.uint8).pin_memory()
            self.pinned[key] = buf
        return buf

    def run(self):
        try:
            torch.cuda.set_device(self.device)
            self.stream = torch.cuda.Stream(device=self.device)
            net, lp = self.make_nets(self.device)
            self.pipe = InferencePipeline(net, lp, self.variant, use_graphs=self.use_graphs)
        except BaseException as exc:            # surfaced by MultiGpuPipeline on the caller's thread
            self.error = exc
            self.ready.set()
            return
        self.ready.set()
        while True:
            job = self.jobs.get()
            if job is None:
                return
            frames, indices, results, done = job
            try:
                self._process(frames, indices, results)
            except BaseException as exc:
                self.error = exc
            done.set()

    @torch.no_grad()
    def _process(self, frames, indices, results):
        shapes = [frames[i].shape[:2] for i in indices]
        with torch.cuda.stream(self.stream):
            for batch in plan_batches(shapes, self.budget):
                ids = [indices[j] for j in batch]
                h, w = frames[ids[0]].shape[:2]
                stage_in = self._pinned("in", (len(ids), h, w, 3))
                for k, i in enumerate(ids):
                    stage_in[k].copy_(torch.from_numpy(np.ascontiguousarray(frames[i])))
                dev_in = stage_in.to(self.device, non_blocking=True)
                out, _ = self.pipe.run_device_graphed(dev_in) if self.use_graphs else self.pipe.run_device(dev_in)
                stage_out = self._pinned("out", (len(ids), h, w, 3))
                stage_out.copy_(out, non_blocking=True)
                self.stream.synchronize()               # this worker's stream only: the other GPUs keep running
                for k, i in enumerate(ids):
                    results[i] = stage_out[k].numpy().copy()


class MultiGpuPipeline:
    """Single-process inference over G GPUs: one worker thread, stream and pinned uint8 staging ring per GPU, frames sharded
    i -> i mod G, mixed frame sizes grouped per shape and micro-batched under a memory budget, no collective anywhere.

        pipe = MultiGpuPipeline(fdn_state_dict, lpnet_state_dict, kind="FDN", devices=[0, 1, 2, 3])
        restored = pipe.run(frames)            # list of uint8 [h,w,3] BGR arrays (any mix of sizes) -> list in the same order
        pipe.run_folder("in/*.png", "out/")    # cv2.imread / cv2.imwrite around run(), like the inference scripts
        pipe.close()

    The library keeps its per-GPU state (FFT twiddle tables, kernel attributes, SM counts) per device, so the workers share one
    process; each worker makes its GPU current for its thread."""

    def __init__(self, fdn_state_dict, lpnet_state_dict, kind="FDN", devices=None, budget_pixels=8 * 640 * 1120, use_graphs=False):
        from . import archs
        if kind not in ("FDN", "FDN_lolv1"):
            raise ValueError("kind must be 'FDN' or 'FDN_lolv1'")
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        if not devices:
            raise RuntimeError("no CUDA device")
        variant = "lolblur" if kind == "FDN" else "lolv1"

        def make_nets(device):
            net = getattr(archs, kind)()
            net.load_state_dict(fdn_state_dict, strict=True)
            lp = archs.I_predict_net()
            lp.load_state_dict(lpnet_state_dict, strict=True)
            return net.to(device).eval(), lp.to(device).eval()

        self.workers = [_Worker(i, "cuda:%d" % d if isinstance(d, int) else d, make_nets, variant, budget_pixels, use_graphs)
                        for i, d in enumerate(devices)]
        for wk in self.workers:
            wk.start()
        for wk in self.workers:
            wk.ready.wait()
            if wk.error is not None:
                raise wk.error

    def run(self, frames):
        frames = list(frames)
        for f in frames:
            if not (isinstance(f, np.ndarray) and f.dtype == np.uint8 and f.ndim == 3 and f.shape[2] == 3):
                raise RuntimeError("frames are uint8 [h,w,3] BGR arrays, as cv2.imread returns them")
        results = [None] * len(frames)
        pending = []
        for wk, indices in zip(self.workers, shard_frames(len(frames), len(self.workers))):
            if indices:
                done = threading.Event()
                wk.jobs.put((frames, indices, results, done))
                pending.append((wk, done))
        for wk, done in pending:
            done.wait()
        for wk, _ in pending:
            if wk.error is not None:
                err, wk.error = wk.error, None
                raise err
        return results

    def run_folder(self, pattern, out_dir, chunk=64):
        """glob -> cv2.imread -> run -> cv2.imwrite (inference_fdn_lolblur.py:42-75), ``chunk`` frames in flight at a time."""
        import glob
        import os
        import cv2
        paths = sorted(glob.glob(pattern))
        os.makedirs(out_dir, exist_ok=True)
        written = []
        for s in range(0, len(paths), chunk):
            part = paths[s:s + chunk]
            outs = self.run([cv2.imread(p, cv2.IMREAD_COLOR) for p in part])
            for p, o in zip(part, outs):
                dst = os.path.join(out_dir, os.path.basename(p))
                cv2.imwrite(dst, o)
                written.append(dst)
        return written

    def close(self):
        for wk in self.workers:
            wk.jobs.put(None)
        for wk in self.workers:
            wk.join(timeout=30)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

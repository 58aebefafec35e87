"""Device-resident version of the reference's inference loop (inference_fdn_lolblur.py:42-75, inference_fdn_lolv1.py:39-68).

The scripts do, per image on the host: cv2.imread -> float32/255 -> img2tensor(bgr2rgb) -> .to(device) -> reflect-pad to a multiple
of 32 -> LPNet -> FDN -> crop -> tensor2img (clamp, *255, round, uint8, RGB->BGR) -> imwrite.  Here only the uint8 frame crosses
PCIe in either direction (4x fewer bytes than fp32) and every step in between is a kernel of libfdn_b200.so:

    pipe = InferencePipeline(net, net_ipred, variant="lolblur")        # or "lolv1"
    out_bgr = pipe(cv2.imread(path))                                   # uint8 [h,w,3] in, uint8 [h,w,3] out

``variant`` selects what the script passes as ``ratio_i``: the LPNet output itself (lolblur, inference_fdn_lolblur.py:65,71) or
mean(gray(img)) / LPNet(img) (lolv1, inference_fdn_lolv1.py:57-64).
"""
import numpy as np
import torch

from . import ops


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
            frames = torch.zeros(b, h, w, 3, dtype=torch.uint8, device=self.device)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):               # warm-up outside capture: twiddle tables, packed weights, kernel attributes
                for _ in range(2):
                    self.run_device(frames)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="relaxed"):    # kernel-attribute calls of the C ABI are not stream work
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

"""SamPredictor over the B200 engines — the drop-in boundary.

Mirrors segment_anything_cs/predictor.py: `set_image :32`, `set_torch_image :72`,
`predict_fg_map :113`, `predict :133`, `predict_torch :214`, `get_image_embedding :294`,
`reset_image :311`, `.device :308`, assignable `.features / .dino_feats / .original_size /
.input_size / .is_image_set / .model`.  Same error behaviour (RuntimeError before set_image,
AssertionError on bad format/shape).  Everything runs under no_grad on the CUDA kernels.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import engine, graphs, ops
from .ops import H16
from .transforms import ResizeLongestSide


class SamPredictor:
    def __init__(self, sam_model, dino_model) -> None:
        self.model = sam_model
        self.dino_model = dino_model
        self.transform = ResizeLongestSide(sam_model.image_encoder.img_size)
        self.reset_image()

    # ------------------------------------------------------------------ image
    def set_image(self, image: np.ndarray, mask: np.ndarray = None, image_format: str = "RGB", cal_image=True):
        assert image_format in ["RGB", "BGR"], f"image_format must be in ['RGB', 'BGR'], is {image_format}."
        if image_format != self.model.image_format:
            image = image[..., ::-1]
        resized = self.transform.apply_image(np.ascontiguousarray(image))
        # interleaved HWC bytes go to the device as they are (one H2D copy); the patch-gather kernel reads HWC
        # blocking copy: `resized` may be a temporary (PIL / cv2 resize output) that is freed as soon as this method
        # returns, and a caller may overwrite its own image buffer right after the call; an asynchronous copy from such
        # memory is only safe through the driver's staging of pageable memory, which is not a contract worth relying on
        # (under compute-sanitizer two otherwise bit-identical runs differed in the last digits).  The stream is idle
        # here anyway: the previous image ended with a device-to-host read of its results.
        hwc = torch.as_tensor(resized).to(self.device)
        mask_t = None
        if mask is not None:
            mask_t = torch.as_tensor(self.transform.apply_image(mask), device=self.device).permute(2, 0, 1).contiguous()[None]
        if cal_image:
            self._set_u8(hwc, tuple(image.shape[:2]), tuple(resized.shape[:2]))
        if mask_t is not None:
            size = self.model.image_encoder.img_size
            h, w = mask_t.shape[-2:]
            return torch.nn.functional.pad(mask_t, (0, size - w, 0, size - h))

    @staticmethod
    def _encode_bind(enc, dino_eng, eng, img_u8: torch.Tensor):
        """Both encoders + the prompt-independent decoder work for one image: the region that is replayed as ONE
        CUDA graph (static shapes, ~380 launches, no host sync).  Returns (features, dino_feats, engine image state)."""
        dino_feats, dino_h = None, None
        if dino_eng is not None and engine.two_streams_enabled() and not graphs.usable():
            # two independent encoders on two streams (engine.two_streams_enabled); the DINOv2 outputs are allocated on
            # the side stream and consumed on this one
            (feats, feat_tok), (dino_f32, dino_h) = engine.interleave_two_streams(
                enc.forward_steps(img_u8), dino_eng.forward_steps(img_u8), img_u8.device)
            cur = torch.cuda.current_stream(img_u8.device)
            for t in (dino_f32, dino_h.hi, dino_h.lo):
                if t is not None:
                    t.record_stream(cur)
            img_u8.record_stream(engine._side_streams[img_u8.device.index])
            dino_feats = dino_f32.view(1, 73, 73, -1)
        else:
            feats, feat_tok = enc.forward(img_u8)
            if dino_eng is not None:
                dino_f32, dino_h = dino_eng.forward(img_u8)
                dino_feats = dino_f32.view(1, 73, 73, -1)
        eng.set_image(feat_tok, dino_h)
        return feats, dino_feats, eng.img

    @torch.no_grad()
    def _set_u8(self, img_u8: torch.Tensor, original_size, input_size):
        """img_u8: uint8 [h,w,3] or [3,h,w] on the device at the model resolution."""
        self.reset_image()
        self.original_size, self.input_size = tuple(original_size), tuple(input_size)
        enc = self.model.image_encoder.engine()
        # without a DINOv2 model: mask / IoU prediction only (SamAutomaticMaskGenerator needs no PWD-Net features)
        dino_eng = self.dino_model.engine() if self.dino_model is not None else None
        eng = self.model.mask_decoder.engine()
        out = None
        if graphs.usable():
            key = ("set_image", tuple(img_u8.shape), id(enc), id(dino_eng))
            g = eng.graphs.get(key)
            if g is None and eng.graphs.second_sight(key):
                g = eng.graphs.put(key, graphs.Graphed(lambda im: SamPredictor._encode_bind(enc, dino_eng, eng, im), [img_u8]))
                g.out[2]["_decode_graphs"] = graphs.GraphCache(3)
            if g is not None:
                feats, dino_feats, state = g(img_u8)
                eng.img = state          # the engine state this replay just refreshed (decode graphs hang off it)
                eng.img_gen += 1
                # the graph's output buffers are rewritten by the next set_image: hand out copies (4 MB + 22 MB),
                # callers may keep `predictor.features` of earlier images (tools/train.py caches them)
                out = (feats.clone(), None if dino_feats is None else dino_feats.clone())
        if out is None:
            feats, dino_feats, _ = self._encode_bind(enc, dino_eng, eng, img_u8)
            out = (feats, dino_feats)
        self.features, self.dino_feats = out
        self._bound = (self.features, self.dino_feats, eng, eng.img_gen)
        self.is_image_set = True

    @torch.no_grad()
    def set_torch_image(self, transformed_image: torch.Tensor, original_image_size: Tuple[int, ...],
                        transformed_mask: torch.Tensor = None, cal_image=True):
        size = self.model.image_encoder.img_size
        assert (len(transformed_image.shape) == 4 and transformed_image.shape[1] == 3
                and max(*transformed_image.shape[2:]) == size), \
            f"set_torch_image input must be BCHW with long side {size}."
        if cal_image:
            img = transformed_image[0]
            if img.dtype != torch.uint8:
                r = img.round()
                if not torch.equal(r, img) or r.min() < 0 or r.max() > 255:
                    raise NotImplementedError("the B200 path takes 8-bit images (as SamPredictor.set_image produces)")
                img = r.to(torch.uint8)
            self._set_u8(img.to(self.device).contiguous(), original_image_size, transformed_image.shape[-2:])
        if transformed_mask is not None:
            h, w = transformed_mask.shape[-2:]
            return torch.nn.functional.pad(transformed_mask, (0, size - w, 0, size - h))

    def _bind(self, feats, dino_feats, feat_tok=None, dino_h=None):
        """Run the prompt-independent decoder work for (features, dino_feats).  When a caller assigned the
        two tensors directly (tools/train.py caches them), derive the engine inputs from them."""
        eng = self.model.mask_decoder.engine()
        if feat_tok is None:
            feat_tok = ops.transpose_f32(feats.reshape(256, 4096).float().contiguous())
        if dino_h is None and dino_feats is not None:
            d = dino_feats.reshape(-1, dino_feats.shape[-1]).float().contiguous()
            _, dino_h, _ = ops.layernorm(d, normalize=False, want_h16=True, split=eng.split)
        eng.set_image(feat_tok, dino_h)
        self._bound = (feats, dino_feats, eng, eng.img_gen)

    def _engine(self):
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        eng = self.model.mask_decoder.engine()
        b = self._bound
        # the decoder engine holds ONE image state; rebind when the caller assigned other embeddings, the engine was
        # rebuilt (weights changed) or another predictor sharing the model set a different image in between
        if (b is None or b[0] is not self.features or b[1] is not self.dino_feats or b[2] is not eng
                or b[3] != eng.img_gen):
            self._bind(self.features, self.dino_feats)
        return self._bound[2]

    # ------------------------------------------------------------------ prompts
    @torch.no_grad()
    def predict_fg_map(self, img_size=None) -> torch.Tensor:
        return self._engine().fg_logits()

    def _coords01(self, point_coords: torch.Tensor, point_labels: torch.Tensor, boxes, mask_input):
        if boxes is not None or mask_input is not None or point_coords is None:
            raise NotImplementedError("only point prompts are on the B200 hot path (boxes / mask inputs are not)")
        if point_coords.dim() != 3 or point_coords.shape[1] != 1:
            raise NotImplementedError("one point per prompt on the B200 hot path")
        # (pt + 0.5) / 1024 in the incoming dtype (float64 from apply_coords), then fp32
        # (prompt_encoder.py:82,215-218)
        c = (point_coords[:, 0, :] + 0.5) / float(self.model.image_encoder.img_size)
        c01 = c.to(torch.float32).contiguous().to(self.device)
        lab = point_labels[:, 0].to(torch.int32).contiguous().to(self.device)
        return c01, lab

    @torch.no_grad()
    def decode_low_res(self, point_coords, point_labels, boxes=None, mask_input=None):
        """low-res logits [P,4,256,256], iou [P,4], cls [P,4,n_class] without the full-size upsample."""
        eng = self._engine()
        c01, lab = self._coords01(point_coords, point_labels, boxes, mask_input)
        return self._decode(eng, c01, lab)[0]

    def _decode(self, eng, c01, lab):
        """-> ((low, iou, cls), from_graph).  The decoder's ~110 launches for a prompt batch are replayed as one CUDA
        graph when the image state came from a set_image graph (static pointers) and this batch size was seen before;
        graph outputs are static buffers, valid until the next decode of the same size."""
        dg = eng.img.get("_decode_graphs") if (eng.img is not None and graphs.usable()) else None
        if dg is not None:
            P = int(c01.shape[0])
            g = dg.get(P)
            if g is None and dg.second_sight(P):
                g = dg.put(P, graphs.Graphed(eng.decode, [c01, lab]))
            if g is not None:
                return g(c01, lab), True
        return eng.decode(c01, lab), False

    @torch.no_grad()
    def predict_torch(self, point_coords: Optional[torch.Tensor], point_labels: Optional[torch.Tensor],
                      boxes: Optional[torch.Tensor] = None, mask_input: Optional[torch.Tensor] = None,
                      multimask_output: bool = True, return_logits: bool = False, attn_sim=None,
                      target_embedding=None):
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        c01, lab = self._coords01(point_coords, point_labels, boxes, mask_input)
        (low, iou, cls), static = self._decode(self._engine(), c01, lab)
        if static:                                       # public API: results must outlive the next call
            low, iou, cls = low.clone(), iou.clone(), cls.clone()
        if not multimask_output:                         # mask_decoder.py:128-134
            low, iou, cls = low[:, :1].contiguous(), iou[:, :1], cls[:, :1]
        P, Cn = low.shape[:2]
        flat = low.reshape(P * Cn, 256, 256)
        thr = float(self.model.mask_threshold)
        m, lg = ops.mask_post_write(flat, None, None, self.input_size, self.original_size, thr,
                                    want_masks=not return_logits, want_logits=return_logits)
        out = lg if return_logits else m
        H, W = self.original_size
        return out.view(P, Cn, H, W), iou, cls, low

    def predict(self, point_coords=None, point_labels=None, box=None, mask_input=None, multimask_output=True,
                return_logits=False, attn_sim=None, target_embedding=None):
        """numpy front end of predict_torch (predictor.py:133-212) for a single prompt.  Returns the reference's four
        values: masks [C,H,W], iou [C], class scores [C,n_class] (numpy) and the low-res logits [C,256,256] (tensor, as
        predictor.py:208-212 leaves it)."""
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        if box is not None or mask_input is not None:
            raise NotImplementedError("only point prompts are on the B200 hot path (boxes / mask inputs are not)")
        assert point_coords is not None and point_labels is not None, "point_labels must be supplied with point_coords."
        pc = self.transform.apply_coords(point_coords, self.original_size)
        ct = torch.as_tensor(pc, dtype=torch.float)[None]
        lt = torch.as_tensor(point_labels, dtype=torch.int)[None]
        masks, iou, cls, low = self.predict_torch(ct, lt, None, None, multimask_output, return_logits=return_logits)
        return masks[0].cpu().numpy(), iou[0].cpu().numpy(), cls[0].cpu().numpy(), low[0]

    # ------------------------------------------------------------------ state
    def get_image_embedding(self) -> torch.Tensor:
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) to generate an embedding.")
        assert self.features is not None, "Features must exist if an image has been set."
        return self.features

    @property
    def device(self) -> torch.device:
        return self.model.device

    def reset_image(self) -> None:
        self.is_image_set = False
        self.features = None
        self.dino_feats = None
        self._bound = None
        self.orig_h = self.orig_w = self.input_h = self.input_w = None

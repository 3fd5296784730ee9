"""What the reference does with the head's output in its test loop (trainval_model.py:216-303), without the
parts that need the backbone or third-party CRF code:

* ``postprocess``      threshold `up` at score_thresh (1e-9, :160), `im_processing.resize_and_crop` to each sample's
                       ground-truth size (util/im_processing.py:25-41) and `compute_mask_IU` (util/eval_tools.py:31-35) --
                       one CUDA kernel over a ragged batch (cmpc_postprocess_iou);
* ``SegEvaluator``     cum_I / cum_U / mean IoU / precision@X bookkeeping and the report text of :267-294;
* ``NpzBatchReader``   the `.npz` batch files of build_batches.py:72-76 read ahead by a thread (util/data_reader.py:8-66).

DenseCRF refinement (:246-259) is pydensecrf, a third-party C++ library, and is not reimplemented.
"""
from __future__ import annotations

import math
import os
import queue
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .parallel import EVAL_SEG_IOU, local_iou_stats, reduce_iou_stats, summarize


def resize_and_crop_meta(im_h: int, im_w: int, out_h: int, out_w: int) -> Tuple[int, int, int, int]:
    """(resized_h, resized_w, crop_h, crop_w) with the integer arithmetic of resize_and_crop (util/im_processing.py:27-32);
    Python's round() is round-half-even like np.round."""
    scale = max(out_h / im_h, out_w / im_w)
    res_h, res_w = int(round(im_h * scale)), int(round(im_w * scale))
    return res_h, res_w, int(math.floor(res_h - out_h) / 2), int(math.floor(res_w - out_w) / 2)


def postprocess(up: torch.Tensor, gt_masks: Sequence, *, score_thresh: float = 1e-9, mode: str = "constant",
                return_masks: bool = True):
    """up: the head's upsampled logits [B, H, W, 1] (or [B, H, W]) on the device; gt_masks: B ground-truth masks [h_b, w_b]
    (numpy / torch, any non-zero = foreground), sizes may differ.  Returns (predicts, I, U): the B resized-and-cropped
    predictions as uint8 device tensors [h_b, w_b] (None unless return_masks) and per-sample int64 intersection / union."""
    if up.device.type != "cuda":
        raise L.CmpcError("postprocess needs the head's output on a CUDA device; there is no CPU path")
    if mode not in ("constant", "reflect"):
        raise L.CmpcError("mode must be 'constant' (skimage <= 0.14 default) or 'reflect'")
    lib = L.lib()
    u = up.reshape(up.shape[0], up.shape[1], up.shape[2]).to(torch.float32).contiguous()
    B, H, W = u.shape
    if len(gt_masks) != B:
        raise L.CmpcError(f"{len(gt_masks)} ground-truth masks for a batch of {B}")
    dev = up.device
    flat, meta, offs, off = [], [], [], 0
    for m in gt_masks:
        m = torch.as_tensor(np.asarray(m) if not torch.is_tensor(m) else m)
        if m.dim() != 2 or m.numel() == 0:
            raise L.CmpcError("each ground-truth mask must be a non-empty 2-D array")
        gh, gw = int(m.shape[0]), int(m.shape[1])
        meta.append((gh, gw) + resize_and_crop_meta(H, W, gh, gw))
        offs.append(off)
        off += gh * gw
        flat.append((m != 0).to(torch.uint8).reshape(-1))
    gt = torch.cat([f.to(dev, non_blocking=True) for f in flat])
    meta_t = torch.tensor(meta, dtype=torch.int32).to(dev)
    offs_t = torch.tensor(offs, dtype=torch.int64).to(dev)
    pred = torch.empty_like(gt) if return_masks else None
    iu = torch.zeros(B, 2, dtype=torch.int64, device=dev)
    L.check(lib.cmpc_postprocess_iou(u.data_ptr(), B, H, W, float(score_thresh), gt.data_ptr(), offs_t.data_ptr(), meta_t.data_ptr(),
                                     1 if mode == "reflect" else 0, pred.data_ptr() if pred is not None else None, iu.data_ptr(),
                                     torch.cuda.current_stream(dev).cuda_stream), "cmpc_postprocess_iou")
    predicts = None
    if return_masks:
        predicts = [pred[o:o + gh * gw].view(gh, gw) for o, (gh, gw, *_r) in zip(offs, meta)]
    return predicts, iu[:, 0], iu[:, 1]


class SegEvaluator:
    """The counters of the reference's test() (trainval_model.py:161-166, 266-274) and its final report (:289-296).  Under
    torch.distributed `finish()` sums the nine numbers over the ranks (one all-reduce)."""

    def __init__(self, device="cpu"):
        self.stats = torch.zeros(9, dtype=torch.float64, device=device)

    def update(self, I: torch.Tensor, U: torch.Tensor) -> None:
        self.stats += local_iou_stats(I.reshape(-1), U.reshape(-1)).to(self.stats.device)

    def finish(self, group=None) -> Dict[str, float]:
        return summarize(reduce_iou_stats(self.stats.clone(), group))

    @staticmethod
    def report(summary: Dict[str, float]) -> str:
        s = "Segmentation evaluation (without DenseCRF):\n"
        for th in EVAL_SEG_IOU:
            s += "precision@%s = %f\n" % (str(th), summary[f"precision@{th}"])
        s += "overall IoU = %f; mean IoU = %f\n" % (summary["overall_iou"], summary["mean_iou"])
        return s


class NpzBatchReader:
    """util/data_reader.py:DataReader: every file of `folder_name` is one batch saved with np.savez (text_batch, im_batch,
    mask_batch, sent_batch -- build_batches.py:72-76 -- plus anything else the file holds, e.g. pre-extracted visual_feat_c3/4/5);
    a daemon thread keeps `prefetch_num` batches loaded; epochs wrap around, reshuffled when shuffle=True."""

    def __init__(self, folder_name: str, prefix: str, shuffle: bool = True, prefetch_num: int = 32, seed: Optional[int] = None):
        self.folder_name, self.prefix, self.shuffle = folder_name, prefix, shuffle
        self.filelist = sorted(f for f in os.listdir(folder_name) if f.endswith(".npz"))
        self.num_batch = len(self.filelist)
        if self.num_batch == 0:
            raise RuntimeError('no batches under %s with prefix "%s"' % (folder_name, prefix))
        self.n_batch = self.n_epoch = 0
        self._rng = np.random.RandomState(seed)
        self._queue: "queue.Queue[Dict[str, np.ndarray]]" = queue.Queue(maxsize=prefetch_num)
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        n, order = 0, np.arange(self.num_batch)
        while True:
            if n == 0 and self.shuffle:
                order = self._rng.permutation(self.num_batch)
            with np.load(os.path.join(self.folder_name, self.filelist[order[n]]), allow_pickle=True) as z:
                batch = dict(z)
            self._queue.put(batch, block=True)
            n = (n + 1) % self.num_batch

    def read_batch(self, is_log: bool = True) -> Dict[str, np.ndarray]:
        if is_log:
            print("data reader: epoch = %d, batch = %d / %d" % (self.n_epoch, self.n_batch, self.num_batch))
        batch = self._queue.get(block=True)
        self.n_batch = (self.n_batch + 1) % self.num_batch
        self.n_epoch += self.n_batch == 0
        return batch

"""The evaluation loop of the reference's ``test()`` (trainval_model.py:216-303) around the device head.

Per sample, as the reference does: read a batch (``util/data_reader.py`` layout: ``text_batch``, ``mask_batch``, ``sent_batch``,
here with the pre-extracted backbone taps ``visual_feat_c3/c4/c5`` instead of ``im_batch`` -- the DeepLab backbone is out of
scope), ``sess.run([pred, up, sigm, words_parse])`` (:232) -> threshold ``up >= score_thresh`` (:243-244) ->
``resize_and_crop`` to the ground-truth size (:245) -> optional DenseCRF (:246-259; pydensecrf is third-party C++ and is NOT
reimplemented: pass ``dcrf=callable(sigm, batch) -> HxW {0,1} mask`` to plug one in) -> ``compute_mask_IU`` (:267) ->
cumulative I / U, mean IoU, precision@{.5..9} (:268-284) -> the report of :287-303.
Under torch.distributed every rank evaluates its contiguous share of the batches and the nine counters are summed with one
all-reduce (parallel.reduce_iou_stats).
"""
from __future__ import annotations

import sys
import time
from typing import Callable, Dict, Optional

import numpy as np
import torch

from .parallel import shard_range
from .postprocess import SegEvaluator, postprocess


def _seq_len(text: np.ndarray) -> int:
    """tokens are right-padded with 0 (util/text_processing.py:55-67 returns the same count as seq_len)"""
    nz = np.nonzero(np.asarray(text).reshape(-1))[0]
    return int(nz[-1]) + 1 if len(nz) else 0


def test(model, reader, *, num_batch: Optional[int] = None, score_thresh: float = 1e-9, dcrf: Optional[Callable] = None,
         resize_mode: str = "constant", group=None, rank: int = 0, world: int = 1, verbose: bool = False, out=sys.stdout,
         iu_fn: Optional[Callable] = None) -> Dict[str, object]:
    """model: the drop-in LSTM_model(batch_size=1, mode='eval'); reader: anything with read_batch(is_log=False) and num_batch
    (postprocess.NpzBatchReader).  Returns {'summary', 'summary_dcrf' (if dcrf), 'IU_result', 'avg_time', 'report'}."""
    dev = model.device
    NN = int(num_batch if num_batch is not None else reader.num_batch)
    lo, hi = shard_range(rank, world, NN)
    ev, ev_d = SegEvaluator(dev), (SegEvaluator(dev) if dcrf is not None else None)
    iu = iu_fn or (lambda up, mask: postprocess(up, [mask], score_thresh=score_thresh, mode=resize_mode, return_masks=False)[1:])
    IU_result, processing_time, seg_total = [], 0.0, 0
    tick = max(NN // 50, 1)
    for n_iter in range(NN):
        batch = reader.read_batch(is_log=False)           # every rank walks the whole reader so that the file order stays aligned
        if not (lo <= n_iter < hi):
            continue
        if rank == 0 and n_iter % tick == 0:                # the reference's progress line (:217-222)
            out.write(str(n_iter // tick // 5) if (n_iter // tick) % 5 == 0 else ".")
            out.flush()
        mask = torch.as_tensor(np.asarray(batch["mask_batch"]).astype(np.float32))
        feed = {k: torch.as_tensor(np.asarray(batch[k]), dtype=torch.float32).to(dev).reshape((1,) + tuple(np.asarray(batch[k]).shape[-3:]))
                for k in ("visual_feat_c3", "visual_feat_c4", "visual_feat_c5")}
        t0 = time.time()
        if "lstm_outputs" in batch:
            lo_t = torch.as_tensor(np.asarray(batch["lstm_outputs"]), dtype=torch.float32).to(dev)
            feed["lstm_outputs"] = lo_t.reshape(1, model.num_steps, -1)
        else:
            text = np.asarray(batch["text_batch"]).reshape(-1)
            sl = int(batch["seq_len"]) if "seq_len" in batch else _seq_len(text)
            feed["lstm_outputs"] = model.encode_words(torch.as_tensor(text.astype(np.int64)).view(1, -1).to(dev),
                                                      torch.tensor([sl], dtype=torch.int32, device=dev))
        scores_val, up_val, sigm_val, words_parse = model.run(["pred", "up", "sigm", "words_parse"], feed_dict=feed)
        if verbose:
            print("Sentence:", batch.get("sent_batch", [""])[0], file=out)
            print("Words type:", words_parse[0][0].cpu().numpy(), file=out)
        I, U = iu(up_val, mask)
        if dcrf is not None:
            pred_dcrf = torch.as_tensor(np.asarray(dcrf(sigm_val[0, :, :, 0].cpu().numpy(), batch)), dtype=torch.float32)
            # a {0,1} mask thresholds to itself at score_thresh; same resize_and_crop + compute_mask_IU as the plain prediction
            Id, Ud = iu(pred_dcrf.reshape(1, model.H, model.W, 1).to(dev), mask)
            ev_d.update(Id, Ud)
        torch.cuda.synchronize(dev) if dev.type == "cuda" else None
        processing_time += time.time() - t0
        ev.update(I, U)
        IU_result.append({"batch_no": n_iter, "I": int(I.reshape(-1)[0]), "U": int(U.reshape(-1)[0])})
        seg_total += 1
    res: Dict[str, object] = {"IU_result": IU_result, "avg_time": processing_time / max(seg_total, 1)}
    res["summary"] = ev.finish(group)
    report = "Avg time: {}\n".format(res["avg_time"]) + SegEvaluator.report(res["summary"])
    if ev_d is not None:
        res["summary_dcrf"] = ev_d.finish(group)
        report += SegEvaluator.report(res["summary_dcrf"]).replace("without DenseCRF", "with DenseCRF")
    res["report"] = report
    if rank == 0:
        print("\n" + report, file=out)
    return res

"""AutoFocus multi-scale inference: the test iterator, `Tester` and the scale pyramid driver.

Host-side mirror of `lib/iterators/MNIteratorTestAutoFocus.py` (:13-141), `lib/inference.py` `Tester.detect` (:100-139),
`Tester.get_detections` (:232-370), `Tester.aggregate` (:152-230, in `inference.aggregate`) and the single-job branch of
`imdb_detection_wrapper` / `detect_scale_worker` (:400-528): same names, argument meaning and returned structures
(`all_boxes[class][image][chip]`, `all_maps[image][chip]`).  What changes is where the arithmetic runs:

  reference                                             here
  cv2 crop / resize / pad in a thread pool              one `sniper_chip_input_hw` launch per batch (uint8 crops over PCIe)
  MXNet module forward (RPN, MultiProposal op on host)  `SniperResNet101.forward_inference` (device MultiProposal)
  bbox_pred / clip / rescale in numpy per chip          the same numpy code (`inference.detect_postprocess`)
  per-(image, class) cpu_soft_nms in a Pool(32)         ONE `sniper_soft_nms_batched` launch for the whole run
  FocusChips with OpenCV                                `chips_inference.add_chips`

roidb entries need `width`, `height`, `image_data` (decoded BGR uint8 [H,W,3]; this image has no OpenCV to decode files)
or an `image_loader`; `inference_crops` is initialised to the full image like the reference does.
"""
import math
import time

import numpy as np
import torch

from . import chips_inference, inference
from ._lib import check, lib


class MNIteratorTestAutoFocus(object):
    """MNIteratorTestAutoFocus(roidb, config, test_scale, batch_size, ...): batches of chips of one test scale, sorted by
    area and grouped by orientation (reset :88-141); a batch is padded to its largest resized chip (_get_batch :36-78)."""

    def __init__(self, roidb, config, test_scale, batch_size=4, threads=8, nGPUs=1, pad_rois_to=400, crop_size=(512, 512),
                 num_classes=None, device="cuda:0", image_loader=None):
        self.roidb, self.cfg, self.test_scale = roidb, config, test_scale
        self.batch_size = batch_size
        self.device = torch.device(device)
        self.image_loader = image_loader
        self.data_name = ['data', 'im_info', 'im_ids', 'chip_ids']
        self.on_gpu = self.device.type == "cuda"        # host-only use (tests of the batching logic): no canvas is built
        self.means = torch.tensor(list(config.network.PIXEL_MEANS), dtype=torch.float32,
                                  device=self.device if self.on_gpu else "cpu")
        self._pix = None
        self.reset()

    def set_scale(self, scale):
        self.test_scale = scale

    def reset(self):
        self.cur_i = 0
        self.crop2im = {}
        widths, heights = [], []
        crop_counter = 0
        for i, r in enumerate(self.roidb):
            local_crop_mapping = {}
            for local_counter, crop in enumerate(r['inference_crops']):
                widths.append(crop[2] - crop[0])
                heights.append(crop[3] - crop[1])
                self.crop2im[crop_counter] = i
                local_crop_mapping[crop_counter] = local_counter
                crop_counter += 1
            r['crop_mapping'] = local_crop_mapping
        self.n_chips = crop_counter
        if crop_counter == 0:
            self.inds, self.size = np.zeros(0, dtype=int), 0
            return
        widths, heights = np.array(widths, dtype=np.float64), np.array(heights, dtype=np.float64)
        order = (widths * heights).argsort()                   # sort based on area
        widths, heights = widths[order], heights[order]
        horz_inds = np.where(widths >= heights)[0]
        vert_inds = np.where(widths < heights)[0]
        bs = self.batch_size
        if horz_inds.shape[0] % bs > 0:
            extra = bs - (horz_inds.shape[0] % bs)
            horz_inds = np.hstack((horz_inds, horz_inds[-extra:]))
        if vert_inds.shape[0] % bs > 0:
            extra = bs - (vert_inds.shape[0] % bs)
            vert_inds = np.hstack((vert_inds, vert_inds[-extra:]))
        inds = np.hstack((horz_inds, vert_inds)).astype(int)
        if inds.shape[0] % bs > 0:                             # a group smaller than its padding request
            extra = bs - (inds.shape[0] % bs)
            inds = np.hstack((inds, np.resize(inds[-extra:], extra)))
        self.inds = order[inds]
        self.size = len(self.inds)

    def __iter__(self):
        return self

    def get_batch_size(self):
        return self.batch_size

    def _image(self, r):
        if 'image_data' in r:
            return r['image_data']
        if self.image_loader is None:
            raise RuntimeError("roidb entry has no 'image_data' and no image_loader was given (no OpenCV in this image)")
        return self.image_loader(r['image'])

    def __next__(self):
        if self.cur_i >= self.size:
            raise StopIteration
        chip_ids = [int(self.inds[i % self.size]) for i in range(self.cur_i, self.cur_i + self.batch_size)]
        im_ids = [self.crop2im[c] for c in chip_ids]
        self.cur_i += self.batch_size
        return self._get_batch([self.roidb[i] for i in im_ids], chip_ids, im_ids)

    next = __next__

    def _get_batch(self, roidb, chip_ids, im_ids):
        n = len(roidb)
        max_size = [0, 0]
        rects, scales, local_chip_ids = [], [], np.zeros(n)
        for i, r in enumerate(roidb):
            scale = chips_inference.image_scale(r['width'], r['height'], self.test_scale)
            cchip_id = r['crop_mapping'][chip_ids[i]]
            cur_chip = r['inference_crops'][cchip_id]
            local_chip_ids[i] = cchip_id
            max_size[0] = max(max_size[0], int(math.ceil((cur_chip[3] - cur_chip[1]) * scale)))
            max_size[1] = max(max_size[1], int(math.ceil((cur_chip[2] - cur_chip[0]) * scale)))
            im = self._image(r)
            y1, y2 = max(int(cur_chip[1]), 0), min(int(cur_chip[3]), im.shape[0])     # worker_autofocus :60-62
            x1, x2 = max(int(cur_chip[0]), 0), min(int(cur_chip[2]), im.shape[1])
            rects.append((im, y1, y2, x1, x2))
            scales.append(scale)
        # the network needs a canvas that survives the stride-16 backbone; the conv kernel wants every layer's output width
        # to be a multiple of 8 pixels (stride 16 -> the canvas width a multiple of 128), heights are free (multiple of 32)
        SH, SW = -(-max_size[0] // 32) * 32, -(-max_size[1] // 128) * 128
        need = sum(max(y2 - y1, 0) * max(x2 - x1, 0) * 3 for _, y1, y2, x1, x2 in rects)
        pix = np.empty(need, dtype=np.uint8)
        table = np.zeros((n, 8), dtype=np.int64)
        im_info = np.zeros((n, 3), dtype=np.float32)
        off = 0
        for k, ((im, y1, y2, x1, x2), scale) in enumerate(zip(rects, scales)):
            h, w = max(y2 - y1, 0), max(x2 - x1, 0)
            m = h * w * 3
            if m:
                pix[off:off + m].reshape(h, w, 3)[...] = im[y1:y2, x1:x2, :3]
            dh, dw = int(np.rint(h * scale)), int(np.rint(w * scale))              # cv2.resize dsize = cvRound(...)
            dh, dw = min(dh, SH), min(dw, SW)
            table[k] = [off, h, w, dh, dw, 0, int(np.float64(scale).view(np.int64)), 0]
            im_info[k] = [dh, dw, scale]
            off += m
        if not self.on_gpu:
            return dict(pixels=pix, table=table, canvas=(SH, SW), im_info=torch.from_numpy(im_info),
                        im_ids=np.array(im_ids, dtype=int), chip_ids=local_chip_ids.astype(int), scales=np.array(scales))
        dev = self.device
        if self._pix is None or self._pix.numel() < need:
            self._pix = torch.empty(int(need * 1.25) + 4096, dtype=torch.uint8, device=dev)
        self._pix[:need].copy_(torch.from_numpy(pix), non_blocking=True)
        tab = torch.from_numpy(table).to(dev)
        data = torch.empty(n, 3, SH, SW, device=dev)
        check(lib().sniper_chip_input_hw(self._pix.data_ptr(), tab.data_ptr(), self.means.data_ptr(), data.data_ptr(), n, SH,
                                         SW, torch.cuda.current_stream().cuda_stream))
        return dict(data=data, im_info=torch.from_numpy(im_info).to(dev), im_ids=np.array(im_ids, dtype=int),
                    chip_ids=local_chip_ids.astype(int), scales=np.array(scales))


class Tester(object):
    """Tester(module, imdb, roidb, test_iter, cfg, ...): `module` = a `SniperResNet101` (forward_inference)."""

    def __init__(self, module, imdb, roidb, test_iter, cfg, rcnn_output_names=None, rpn_output_names=None, logger=None,
                 batch_size=None, num_classes=None):
        self.net, self.imdb, self.roidb, self.test_iter, self.cfg = module, imdb, roidb, test_iter, cfg
        self.batch_size = batch_size
        self.num_images = len(roidb)
        self.num_classes = num_classes or (imdb.num_classes if imdb is not None else cfg.dataset.NUM_CLASSES)
        self.detect_time = self.post_time = 0.0

    def detect(self, batch, scales=None):
        """(scores, boxes, data, im_ids, maps, chip_ids) of one batch (:100-139): scores [R,K] and boxes [R,4] per chip,
        boxes in ORIGINAL-image scale relative to the chip origin."""
        want_map = bool(getattr(self.cfg.TEST, "AUTO_FOCUS", False)) and self.net.af is not None
        out = self.net.forward_inference(batch['data'], batch['im_info'], autofocus=want_map)
        rois, _, cls_prob, deltas = out[:4]
        B = batch['data'].shape[0]
        scores, preds = inference.detect_postprocess(rois.cpu().numpy(), cls_prob.cpu().numpy(), deltas.cpu().numpy(),
                                                     batch['im_info'].cpu().numpy(), B)
        maps = []
        if want_map:
            fm = out[4].cpu().numpy()
            info = batch['im_info'].cpu().numpy()
            for idx in range(B):       # the map of the chip itself (the canvas is padded to the batch maximum)
                mh, mw = int(math.ceil(info[idx, 0] / 16.0)), int(math.ceil(info[idx, 1] / 16.0))
                maps.append(fm[idx, :mh, :mw].copy())
        return scores, preds, batch, batch['im_ids'], maps, batch['chip_ids']

    def get_detections(self, cls_thresh=1e-3, cache_name='cache', evaluate=False, vis=False, vis_path=None,
                       do_pruning=False, autofocus=False, vis_ext='.png'):
        """(all_boxes, all_maps) (:232-370): all_boxes[j][im][chip] = [n,5], all_maps[im][chip] = (None, FocusPixel map)."""
        n_chips = [len(r['inference_crops']) for r in self.roidb]
        all_boxes = [[[np.zeros((0, 5), np.float32) for _ in range(n_chips[i])] for i in range(self.num_images)]
                     for _ in range(self.num_classes)]
        all_maps = [[[] for _ in range(n_chips[i])] for i in range(self.num_images)]
        for batch in self.test_iter:
            t0 = time.time()
            scores, boxes, _, im_ids, maps, chip_ids = self.detect(batch)
            torch.cuda.synchronize()
            self.detect_time += time.time() - t0
            t0 = time.time()
            for i, (cscores, cboxes, im_id, chip_id) in enumerate(zip(scores, boxes, im_ids, chip_ids)):
                if autofocus and maps:
                    all_maps[im_id][chip_id] = (None, maps[i])
                # score threshold per class; with do_pruning also: project back to the image and drop detections cut
                # by a chip border (:335-351), for all classes in one pass
                r = self.roidb[im_id]
                prune = (r['inference_crops'][chip_id], r['width'], r['height']) if do_pruning else None
                dets = inference.threshold_detections(cscores, cboxes, self.num_classes, cls_thresh, prune=prune)
                for j in range(1, self.num_classes):
                    all_boxes[j][im_id][chip_id] = dets[j]
            self.post_time += time.time() - t0
        return all_boxes, all_maps

    def aggregate(self, scale_cls_dets, vis=False, cache_name='cache', vis_path=None, vis_name=None, pre_nms_db_divide=10,
                  vis_ext='.png', backend="device"):
        t = self.cfg.TEST
        return inference.aggregate(scale_cls_dets, t.VALID_RANGES, self.num_images, self.num_classes,
                                   sigma=getattr(t, "NMS_SIGMA", 0.55), nms_thresh=getattr(t, "NMS", 0.3),
                                   max_per_image=t.MAX_PER_IMAGE, backend=backend,
                                   device=str(self.net.device) if self.net is not None else "cuda")


def imdb_detection_wrapper(net, config, roidb, num_classes=None, batch_images=None, device="cuda:0", nms_backend="device"):
    """The CONCURRENT_JOBS == 1 path of imdb_detection_wrapper (:436-528) without the dataset evaluation: every scale of
    TEST.SCALES in turn (FocusChips of scale i feed scale i + 1 when DO_PRUNING[i + 1]), then the cross-scale aggregation.
    Returns (all_boxes[class][image], stats) with per-scale chip counts / detect / post-processing seconds, the NMS
    latency and the pixel fraction processed."""
    T = config.TEST
    for r in roidb:
        r['inference_crops'] = np.array([[0, 0, r['width'], r['height']]])
    detections, stats = [], dict(scales=[])
    nb = batch_images or list(T.BATCH_IMAGES)
    for scale_i, (nbatch, scale) in enumerate(zip(nb, T.SCALES)):
        it = MNIteratorTestAutoFocus(roidb, config, scale, batch_size=nbatch, device=device)
        tester = Tester(net, None, roidb, it, config, batch_size=nbatch, num_classes=num_classes)
        t0 = time.time()
        dets, maps = tester.get_detections(do_pruning=bool(T.DO_PRUNING[scale_i]), autofocus=bool(T.AUTO_FOCUS))
        st = dict(scale=tuple(scale), chips=int(it.n_chips), batches=int(it.size // max(nbatch, 1)),
                  detect_s=tester.detect_time, post_s=tester.post_time, total_s=time.time() - t0)
        if scale_i + 1 < len(T.SCALES) and T.DO_PRUNING[scale_i + 1] and T.AUTO_FOCUS:
            t1 = time.time()
            area = chips_inference.add_chips(roidb, maps, scale_i, config)
            st.update(chip_gen_s=time.time() - t1, pixels_next_scale_pct=100.0 * area[0] / max(area[1], 1e-12))
        # the per-chip lists of a scale become one list per image for the aggregation
        detections.append(dets)
        stats['scales'].append(st)
    tester = Tester(net, None, roidb, None, config, num_classes=num_classes)
    t0 = time.time()
    all_boxes = tester.aggregate(detections, backend=nms_backend)
    torch.cuda.synchronize()
    stats['nms_s'] = time.time() - t0
    return all_boxes, stats

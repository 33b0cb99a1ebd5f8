"""`OCR` - the reference's user-facing class on the B200 path, same constructor and method arguments.

Mirrors /root/reference/src/ocrfa/ocr.py: `OCR.__init__` (:44-202), `transform_ua` (:283-355), `recognize`
(:357-502), `readtext` (:504-582), `readtext_batched` (:684-789), and the result helpers of utils.py it calls
(`reformat_input` :1525, `get_paragraph` :1442, `set_result_with_confidence` :1615, `merge_to_free` :1176, `diff`
:1438). The page stays on the GPU from the moment it is uploaded: grey conversion (ocrfa_bgr2gray_u8), detector input
(ocrfa_det_preprocess_u8), CRAFT forward, component labeling (ocrfa_det_components), crops (ocrfa_crop_resize_u8 /
ocrfa_warp_perspective_u8), AlignCollate, recognizer forward and decode all run in libocrfa_b200.so; this module is the
list / string logic between them. There is no CPU path: `gpu=False` raises.

Differences from the reference, all explicit:
  * no downloads (this package never touches the network): a missing model file raises FileNotFoundError, as the
    reference does with download_enabled=False; `verify_md5=False` accepts weights other than the released files;
  * DBNet (`detect_network="dbnet18"`), `optimal_num_chars` and `rotation_info` angles that are not multiples of 90
    degrees (scipy's spline resampling) are not offered (NotImplementedError);
  * right-to-left display order (ocr.py:474-478) uses python-bidi's `get_display` when that package is importable and
    `ocrfa_b200.bidi.get_display` - a restatement of the algorithm it implements, UAX #9 - otherwise
    (`OCR.bidi_backend` names the one in use); `display_order=False` returns logical-order strings.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import json
import os
from logging import getLogger

import numpy as np
import torch

from . import _lib
from .charset import ARABIC_G1_CHARACTERS, ARABIC_G1_SYMBOLS, FA_LANG_LETTERS
from .craft import get_detector
from .crops import DeviceLine, get_image_list
from .detection import get_textbox, group_text_box
from .recognition import get_recognizer, get_text

LOGGER = getLogger(__name__)

# utils.py:357-359, 361-378, 408-416: where the reference keeps its model files, and the two it loads by default
MODULE_PATH = os.environ.get("OCR_FA_MODULE_PATH") or os.environ.get("MODULE_PATH") or os.path.expanduser("~/.OCR-FA/")
DETECTION_MODELS = {"craft": {"filename": "craft_mlt_25k.pth", "md5sum": "2f8227d2def4037cdb3b34389dcf9ec1"}}
RECOGNITION_MODELS = {"arabic_g1": {"filename": "arabic.pth", "model_script": "arabic",
                                    "md5sum": "993074555550e4e06a6077d55ff0449a", "symbols": ARABIC_G1_SYMBOLS,
                                    "characters": ARABIC_G1_CHARACTERS}}

try:  # ocr.py:30
    from bidi.algorithm import get_display as _bidi_get_display
    _BIDI_BACKEND = "python-bidi"
except Exception:  # python-bidi is not part of this image: the restatement of the algorithm it implements (UAX #9)
    from .bidi import get_display as _bidi_get_display
    _BIDI_BACKEND = "ocrfa_b200.bidi"


def calculate_md5(fname) -> str:
    """utils.py:1430-1435."""
    digest = hashlib.md5()
    with open(fname, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 16), b""):
            digest.update(chunk)
    return digest.hexdigest()


def diff(input_list):
    """utils.py:1438-1439."""
    return max(input_list) - min(input_list)


def bgr2gray_device(img: torch.Tensor) -> torch.Tensor:
    """cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) for a uint8 [H, W, 3] CUDA tensor, bit-exact with OpenCV."""
    if img.dim() != 3 or img.shape[2] != 3 or img.dtype != torch.uint8 or not img.is_cuda:
        raise ValueError("img must be a uint8 [H, W, 3] CUDA tensor")
    if img.stride(2) != 1 or img.stride(1) != 3:
        img = img.contiguous()
    h, w = int(img.shape[0]), int(img.shape[1])
    with torch.cuda.device(img.device):
        gray = torch.empty((h, w), dtype=torch.uint8, device=img.device)
        _lib.check(_lib.load().ocrfa_bgr2gray_u8(C.c_void_p(img.data_ptr()), int(img.stride(0)), h, w,
                                                 C.c_void_p(gray.data_ptr()), w,
                                                 C.c_void_p(torch.cuda.current_stream(img.device).cuda_stream)))
    return gray


def reformat_input(image, device="cuda"):
    """utils.py:1525-1565: (img [H, W, 3], img_cv_grey [H, W]), both uint8 CUDA tensors. Files and byte strings are
    decoded on the host by OpenCV (I/O); array inputs are uploaded once and converted on the GPU."""
    if not isinstance(image, (str, bytes, torch.Tensor, np.ndarray)) and not hasattr(image, "__array__"):
        raise ValueError("Invalid input type. Supporting format = string(file path), bytes, numpy array")
    if isinstance(image, str) and (image.startswith("http://") or image.startswith("https://")):
        raise ValueError("ocrfa_b200 does not fetch URLs: pass a file path, bytes or an array")
    if isinstance(image, (torch.Tensor, np.ndarray)) and image.dtype not in (torch.uint8, np.uint8):
        raise ValueError("image arrays must be uint8")
    if isinstance(image, str) and not os.path.exists(os.path.expanduser(image)):
        raise ValueError(f"cannot read image file {image!r}")
    dev = _lib.resolve_device(device)
    if isinstance(image, (str, bytes)):
        import cv2

        if isinstance(image, str):
            if image.startswith("http://") or image.startswith("https://"):
                raise ValueError("ocrfa_b200 does not fetch URLs: pass a file path, bytes or an array")
            path = os.path.expanduser(image)
            grey = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
            colour = cv2.imread(path, cv2.IMREAD_COLOR)
            if grey is None or colour is None:
                raise ValueError(f"cannot read image file {image!r}")
            colour = cv2.cvtColor(colour, cv2.COLOR_BGR2RGB)  # loadImage returns RGB (utils.py:466-476)
            return torch.from_numpy(colour).to(dev), torch.from_numpy(grey).to(dev)
        colour = cv2.imdecode(np.frombuffer(image, np.uint8), cv2.IMREAD_COLOR)
        if colour is None:
            raise ValueError("cannot decode the image bytes")
        img = torch.from_numpy(cv2.cvtColor(colour, cv2.COLOR_BGR2RGB)).to(dev)
        return img, bgr2gray_device(img)
    if isinstance(image, torch.Tensor):
        arr = image if image.is_cuda else image.to(dev)
    elif isinstance(image, np.ndarray):
        arr = torch.from_numpy(np.ascontiguousarray(image)).to(dev)
    elif hasattr(image, "__array__"):  # PIL images (the reference names JpegImageFile): RGB -> BGR first
        arr = torch.from_numpy(np.ascontiguousarray(np.array(image))).to(dev)
        if arr.dim() == 3 and arr.shape[2] == 3:
            arr = arr.flip(2).contiguous()
            return arr, bgr2gray_device(arr)
    else:
        raise ValueError("Invalid input type. Supporting format = string(file path), bytes, numpy array")
    if arr.dtype != torch.uint8:
        raise ValueError("image arrays must be uint8")
    if arr.dim() == 2 or (arr.dim() == 3 and arr.shape[2] == 1):
        grey = arr if arr.dim() == 2 else arr[:, :, 0]
        grey = grey.contiguous()
        return grey.unsqueeze(2).expand(-1, -1, 3).contiguous(), grey  # cv2.COLOR_GRAY2BGR replicates the channel
    if arr.dim() == 3 and arr.shape[2] == 3:
        return arr, bgr2gray_device(arr)
    if arr.dim() == 3 and arr.shape[2] == 4:
        img = arr[:, :, :3].flip(2).contiguous()  # drop alpha, RGB -> BGR
        return img, bgr2gray_device(img)
    raise ValueError()


def reformat_input_batched(image, n_width=None, n_height=None, device="cuda"):
    """utils.py:1568-1596 for lists / 4-D arrays of equally sized images: (list of img, list of grey)."""
    if (isinstance(image, np.ndarray) and image.ndim == 4) or isinstance(image, list):
        if n_width is not None or n_height is not None:
            raise NotImplementedError("readtext_batched: resize the images to one shape before calling")
        pairs = [reformat_input(single, device) for single in image]
        if len({tuple(p[0].shape) for p in pairs}) > 1:
            raise ValueError("The input image array contains images of different sizes. "
                             "Please resize all images to same shape or pass n_width, n_height to auto-resize")
        return [p[0] for p in pairs], [p[1] for p in pairs]
    img, grey = reformat_input(image, device)
    return [img], [grey]


def make_rotated_img_list(rotationInfo, img_list):
    """utils.py:1599-1612 for right angles: appends, per angle, every crop rotated by it. The reference calls
    scipy.ndimage.rotate(img, angle, reshape=True), which for multiples of 90 degrees is the exact permutation
    np.rot90(img, angle / 90); those run on the GPU (ocrfa_rot90_u8) on the device-resident crops. Other angles need
    scipy's cubic-spline resampling, which this package does not restate."""
    turns = []
    for angle in rotationInfo:
        if angle % 90 != 0:
            raise NotImplementedError("rotation_info supports multiples of 90 degrees only")
        turns.append(int(angle // 90) % 4)
    if not img_list:
        return img_list[:]
    lib = _lib.load()
    sizes = [((line.shape[0] * line.shape[1] + 3) & ~3) for _, line in img_list]
    dev = img_list[0][1].buffer.device
    result = []
    with torch.cuda.device(dev):
        # originals and rotations share ONE packed buffer (a batch handed to AlignCollate reads from a single buffer)
        out = torch.empty(max(4, sum(sizes) * (1 + len(turns))), dtype=torch.uint8, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        off = 0
        for k in [0] + turns:
            for (box, line), nbytes in zip(img_list, sizes):
                h, w = line.shape
                if k == 0:
                    out[off:off + h * w].copy_(line.buffer[line.off:line.off + h * w])
                    rotated = DeviceLine(out, off, h, w)
                else:
                    _lib.check(lib.ocrfa_rot90_u8(C.c_void_p(line.buffer.data_ptr() + line.off), h, w, k,
                                                  C.c_void_p(out.data_ptr() + off), stream))
                    rotated = DeviceLine(out, off, *((w, h) if k & 1 else (h, w)))
                result.append((box, rotated))
                off += nbytes
    return result


def set_result_with_confidence(results):
    """utils.py:1615-1631: per image, the augmentation with the highest confidence (first one on ties)."""
    final = []
    for col in range(len(results[0])):
        best_row = max(((row, results[row][col][2]) for row in range(len(results))), key=lambda rc: rc[1])[0]
        final.append(results[best_row][col])
    return final


def get_paragraph(raw_result, x_ths=1, y_ths=0.5, mode="ltr"):
    """utils.py:1442-1509: greedy clustering of boxes into paragraphs, then reading order inside each."""
    group = []  # [text, min_x, max_x, min_y, max_y, height, y_centre, paragraph id (0 = unassigned)]
    for box in raw_result:
        xs = [int(pt[0]) for pt in box[0]]
        ys = [int(pt[1]) for pt in box[0]]
        group.append([box[1], min(xs), max(xs), min(ys), max(ys), max(ys) - min(ys), 0.5 * (min(ys) + max(ys)), 0])
    current = 1
    while any(b[7] == 0 for b in group):
        unassigned = [b for b in group if b[7] == 0]
        members = [b for b in group if b[7] == current]
        if not members:
            unassigned[0][7] = current
            continue
        mean_height = np.mean([b[5] for b in members])
        x_lo = min(b[1] for b in members) - x_ths * mean_height
        x_hi = max(b[2] for b in members) + x_ths * mean_height
        y_lo = min(b[3] for b in members) - y_ths * mean_height
        y_hi = max(b[4] for b in members) + y_ths * mean_height
        for b in unassigned:
            if ((x_lo <= b[1] <= x_hi) or (x_lo <= b[2] <= x_hi)) and ((y_lo <= b[3] <= y_hi) or (y_lo <= b[4] <= y_hi)):
                b[7] = current
                break
        else:
            current += 1
    result = []
    for gid in set(b[7] for b in group):
        members = [b for b in group if b[7] == gid]
        mean_height = np.mean([b[5] for b in members])
        x_lo, x_hi = min(b[1] for b in members), max(b[2] for b in members)
        y_lo, y_hi = min(b[3] for b in members), max(b[4] for b in members)
        text = ""
        while members:
            top = min(b[6] for b in members)
            row = [b for b in members if b[6] < top + 0.4 * mean_height]
            if mode == "ltr":
                edge = min(b[1] for b in row)
                pick = [b for b in row if b[1] == edge][-1]
            elif mode == "rtl":
                edge = max(b[2] for b in row)
                pick = [b for b in row if b[2] == edge][-1]
            text += " " + pick[0]
            members.remove(pick)
        result.append([[[x_lo, y_lo], [x_hi, y_lo], [x_hi, y_hi], [x_lo, y_hi]], text[1:]])
    return result


def merge_to_free(merge_result, free_list):
    """utils.py:1176-1215: re-inserts the results of the slanted boxes (they come last) at their reading position
    among the rows of horizontal results."""
    if not free_list:
        return merge_result
    free_results = merge_result[-len(free_list):]
    level_results = merge_result[:-len(free_list)]
    rows, row = [], []
    for idx, r in enumerate(level_results):
        if idx == len(level_results) - 1:
            row.append(r)
            rows.append(row)
            row = []
        elif row == [] or row[-1][0] < r[0]:
            row.append(r)
        else:
            rows.append(row)
            row = [r]
    for item in free_results:
        y_pos = len(rows)
        x_pos = len(rows[y_pos - 1])
        for i, candidate in enumerate(rows[1:]):
            if item[0][0][1] < candidate[0][0][0][1]:
                y_pos = i
                break
        for i, candidate in enumerate(rows[y_pos]):
            if item[0][0][0] < candidate[0][0][0]:
                x_pos = i
                break
        rows[y_pos].insert(x_pos, item)
    merged = []
    for r in rows:
        merged.extend(r)
    return merged


class OCR:
    """ocr.py:42. Farsi reader: CRAFT detector + generation-1 recognizer (ResNet-29 + 2 x BiLSTM + CTC) on one B200."""

    def __init__(self, gpu=True, recog_network="standard", download_enabled=True, detector=True, recognizer=True,
                 model_storage_directory=None, user_network_directory=None, detect_network="craft", verbose=True,
                 quantize=True, cudnn_benchmark=False, *, precision="f16x3", detector_precision="f16x3", verify_md5=True,
                 display_order=True):
        self.verbose = verbose
        self.download_enabled = download_enabled
        self.model_storage_directory = model_storage_directory or (MODULE_PATH + "/model")
        self.user_network_directory = user_network_directory or (MODULE_PATH + "/user_network")
        if gpu is False or str(gpu) == "cpu":
            raise _lib.OcrfaError("ocrfa_b200 has no CPU path: OCR(gpu=False) is not available")
        if not torch.cuda.is_available():
            raise _lib.OcrfaError("ocrfa_b200 needs a CUDA device (B200, sm_100a); none is visible")
        self.device = "cuda" if gpu is True else gpu
        self.quantize = quantize,  # the reference stores a 1-tuple here (ocr.py:96); kept for attribute parity
        self.cudnn_benchmark = cudnn_benchmark
        self.precision, self.detector_precision, self.verify_md5 = precision, detector_precision, verify_md5
        self.detection_models = DETECTION_MODELS
        self.recognition_models = {"gen1": RECOGNITION_MODELS, "gen2": {}}
        self.support_detection_network = ["craft"]
        self.bidi_available = True
        self.bidi_backend = _BIDI_BACKEND  # which get_display runs: the reference's dependency, or ocrfa_b200.bidi
        self.display_order = bool(display_order)
        if recog_network not in ("standard", "arabic_g1"):
            raise ValueError("only the reference's default Farsi recognizer ('standard' / 'arabic_g1') is available")
        # language_list is fixed to ["fa"] (ocr.py:57): setModelLanguage("arabic", ...), arabic_g1, generation1
        model = RECOGNITION_MODELS["arabic_g1"]
        self.model_lang = model["model_script"]
        self.character = model["characters"]
        self.setLanguageList(["fa"], model)
        if detector:
            self.detector = self.initDetector(self.getDetectorPath(detect_network))
        if recognizer:
            model_path = self._checked_path(model)
            self.recognizer, self.converter = get_recognizer(
                "generation1", {"input_channel": 1, "output_channel": 512, "hidden_size": 512}, self.character, {},
                {"fa": os.path.join(os.path.dirname(__file__), "dict", "fa.txt")}, model_path, device=self.device,
                quantize=quantize, precision=precision)

    # -- model files ---------------------------------------------------------------------------------------------
    def _checked_path(self, entry) -> str:
        path = os.path.join(self.model_storage_directory, entry["filename"])
        if not os.path.isfile(path):
            raise FileNotFoundError("Missing %s and downloads disabled" % path)
        if self.verify_md5 and calculate_md5(path) != entry["md5sum"]:
            raise FileNotFoundError("MD5 mismatch for %s and downloads disabled" % path)
        return path

    def getDetectorPath(self, detect_network):
        """ocr.py:204-245 without the download branches."""
        if detect_network in ("dbnet18",):
            raise NotImplementedError("the DBNet detector is not part of ocrfa_b200; use detect_network='craft'")
        if detect_network not in self.support_detection_network:
            raise RuntimeError("Unsupport detector network. Support networks are {}.".format(
                ", ".join(self.support_detection_network)))
        self.detect_network = detect_network
        self.get_textbox = get_textbox
        self.get_detector = get_detector
        return self._checked_path(self.detection_models[detect_network])

    def initDetector(self, detector_path):
        return self.get_detector(detector_path, device=self.device, quantize=self.quantize,
                                 cudnn_benchmark=self.cudnn_benchmark, precision=self.detector_precision)

    def setDetector(self, detect_network):
        self.detector = self.initDetector(self.getDetectorPath(detect_network))

    def setLanguageList(self, language_list, model):
        """ocr.py:273-297: the Farsi letters united with the model's symbols."""
        if model.get("symbols"):
            symbol = model["symbols"]
        elif model.get("character_list"):
            symbol = model["character_list"]
        else:
            symbol = "0123456789!"
        self.lang_char = "".join(set(FA_LANG_LETTERS).union(set(symbol)))

    # -- detection ---------------------------------------------------------------------------------------------------
    def transform_ua(self, img, min_size=20, text_threshold=0.7, low_text=0.4, link_threshold=0.4, canvas_size=2560,
                     mag_ratio=1., slope_ths=0.1, ycenter_ths=0.5, height_ths=0.5, width_ths=0.5, add_margin=0.1,
                     reformat=True, optimal_num_chars=None, threshold=0.2, bbox_min_score=0.2, bbox_min_size=3,
                     max_candidates=0):
        """ocr.py:299-355: (horizontal_list_agg, free_list_agg), one entry per image."""
        if reformat:
            img, _ = reformat_input(img, self.device)
        if isinstance(img, list):
            img = torch.stack(img)
        text_box_list = self.get_textbox(self.detector, img, canvas_size=canvas_size, mag_ratio=mag_ratio,
                                         text_threshold=text_threshold, link_threshold=link_threshold, low_text=low_text,
                                         poly=False, device=self.device, optimal_num_chars=optimal_num_chars,
                                         threshold=threshold, bbox_min_score=bbox_min_score, bbox_min_size=bbox_min_size,
                                         max_candidates=max_candidates)
        horizontal_list_agg, free_list_agg = [], []
        for text_box in text_box_list:
            horizontal_list, free_list = group_text_box(text_box, slope_ths, ycenter_ths, height_ths, width_ths,
                                                        add_margin, (optimal_num_chars is None))
            if min_size:
                horizontal_list = [b for b in horizontal_list if max(b[1] - b[0], b[3] - b[2]) > min_size]
                free_list = [b for b in free_list if max(diff([c[0] for c in b]), diff([c[1] for c in b])) > min_size]
            horizontal_list_agg.append(horizontal_list)
            free_list_agg.append(free_list)
        return horizontal_list_agg, free_list_agg

    # -- recognition -------------------------------------------------------------------------------------------------
    def _read_boxes(self, page, h_list, f_list, img_height, ignore_char, decoder, beamWidth, batch_size, contrast_ths,
                    adjust_contrast, filter_ths, workers):
        image_list, max_width = get_image_list(h_list, f_list, page, model_height=img_height, as_device=True)
        return get_text(self.character, img_height, int(max_width), self.recognizer, self.converter, image_list,
                        ignore_char, decoder, beamWidth, batch_size, contrast_ths, adjust_contrast, filter_ths, workers,
                        self.device)

    def recognize(self, img_cv_grey, horizontal_list=None, free_list=None, decoder="greedy", beamWidth=5, batch_size=1,
                  workers=0, allowlist=None, blocklist=None, detail=1, rotation_info=None, paragraph=False,
                  contrast_ths=0.1, adjust_contrast=0.5, filter_ths=0.003, y_ths=0.5, x_ths=1.0, reformat=True,
                  output_format="standard", img_height=64):
        """ocr.py:357-502. `img_cv_grey`: the grey page (uint8 [H, W], numpy or CUDA tensor) or, with reformat=True,
        anything `reformat_input` accepts."""
        if reformat:
            _, img_cv_grey = reformat_input(img_cv_grey, self.device)
        elif isinstance(img_cv_grey, np.ndarray):
            img_cv_grey = torch.from_numpy(np.ascontiguousarray(img_cv_grey)).to(_lib.resolve_device(self.device))
        if allowlist:
            ignore_char = "".join(set(self.character) - set(allowlist))
        elif blocklist:
            ignore_char = "".join(set(blocklist))
        else:
            ignore_char = "".join(set(self.character) - set(self.lang_char))
        if horizontal_list is None and free_list is None:
            y_max, x_max = img_cv_grey.shape
            horizontal_list = [[0, x_max, 0, y_max]]
            free_list = []
        args = (img_height, ignore_char, decoder, beamWidth, batch_size, contrast_ths, adjust_contrast, filter_ths, workers)
        if batch_size == 1 and not rotation_info:
            # the reference reads box by box here, each line padded to ITS OWN width (ocr.py:398-444); batching
            # (batch_size > 1) pads every line to the widest one, which changes the logits - so the split is kept
            result = []
            for bbox in horizontal_list:
                result += self._read_boxes(img_cv_grey, [bbox], [], *args)
            for bbox in free_list:
                result += self._read_boxes(img_cv_grey, [], [bbox], *args)
        elif not rotation_info:
            result = self._read_boxes(img_cv_grey, horizontal_list, free_list, *args)
        else:
            # test-time augmentation (ocr.py:446-473): every crop is also read rotated, the most confident reading wins
            image_list, max_width = get_image_list(horizontal_list, free_list, img_cv_grey, model_height=img_height,
                                                   as_device=True)
            image_len = len(image_list)
            if image_list:
                image_list = make_rotated_img_list(rotation_info, image_list)
                max_width = max(max_width, img_height)
            result = get_text(self.character, img_height, int(max_width), self.recognizer, self.converter, image_list,
                              ignore_char, decoder, beamWidth, batch_size, contrast_ths, adjust_contrast, filter_ths,
                              workers, self.device)
            if horizontal_list + free_list:
                result = set_result_with_confidence(
                    [result[image_len * i:image_len * (i + 1)] for i in range(len(rotation_info) + 1)])
        if self.model_lang == "arabic":
            direction_mode = "rtl"
            result = [list(item) for item in result]
            if self.display_order:  # ocr.py:474-478
                for item in result:
                    item[1] = _bidi_get_display(item[1])
        else:
            direction_mode = "ltr"
        if paragraph:
            result = get_paragraph(result, x_ths=x_ths, y_ths=y_ths, mode=direction_mode)
        if detail == 0:
            return [item[1] for item in result]
        if output_format == "dict":
            if paragraph:
                return [{"boxes": item[0], "text": item[1]} for item in result]
            return [{"boxes": item[0], "text": item[1], "confident": item[2]} for item in result]
        if output_format == "json":
            rows = []
            for item in result:
                row = {"boxes": [list(map(int, pts)) for pts in item[0]], "text": item[1]}
                if not paragraph:
                    row["confident"] = item[2]
                rows.append(json.dumps(row, ensure_ascii=False))
            return rows
        if output_format == "free_merge":
            return merge_to_free(result, free_list)
        return result

    def readtext(self, image, decoder="greedy", beamWidth=5, batch_size=1, workers=0, allowlist=None, blocklist=None,
                 detail=1, rotation_info=None, paragraph=False, min_size=20, contrast_ths=0.1, adjust_contrast=0.5,
                 filter_ths=0.003, text_threshold=0.7, low_text=0.4, link_threshold=0.4, canvas_size=2560, mag_ratio=1.,
                 slope_ths=0.1, ycenter_ths=0.5, height_ths=0.5, width_ths=0.5, y_ths=0.5, x_ths=1.0, add_margin=0.1,
                 threshold=0.2, bbox_min_score=0.2, bbox_min_size=3, max_candidates=0, output_format="standard"):
        """ocr.py:504-582: detect, group, crop, recognize. One upload of the page; everything between runs on the GPU."""
        img, img_cv_grey = reformat_input(image, self.device)
        horizontal_list, free_list = self.transform_ua(
            img, min_size=min_size, text_threshold=text_threshold, low_text=low_text, link_threshold=link_threshold,
            canvas_size=canvas_size, mag_ratio=mag_ratio, slope_ths=slope_ths, ycenter_ths=ycenter_ths,
            height_ths=height_ths, width_ths=width_ths, add_margin=add_margin, reformat=False, threshold=threshold,
            bbox_min_score=bbox_min_score, bbox_min_size=bbox_min_size, max_candidates=max_candidates)
        horizontal_list, free_list = horizontal_list[0], free_list[0]
        return self.recognize(img_cv_grey, horizontal_list, free_list, decoder, beamWidth, batch_size, workers, allowlist,
                              blocklist, detail, rotation_info, paragraph, contrast_ths, adjust_contrast, filter_ths,
                              y_ths, x_ths, False, output_format)

    def readtext_batched(self, image, n_width=None, n_height=None, decoder="greedy", beamWidth=5, batch_size=1, workers=0,
                         allowlist=None, blocklist=None, detail=1, rotation_info=None, paragraph=False, min_size=20,
                         contrast_ths=0.1, adjust_contrast=0.5, filter_ths=0.003, text_threshold=0.7, low_text=0.4,
                         link_threshold=0.4, canvas_size=2560, mag_ratio=1., slope_ths=0.1, ycenter_ths=0.5,
                         height_ths=0.5, width_ths=0.5, y_ths=0.5, x_ths=1.0, add_margin=0.1, threshold=0.2,
                         bbox_min_score=0.2, bbox_min_size=3, max_candidates=0, output_format="standard"):
        """ocr.py:684-789: a list (or 4-D array) of equally sized pages; one detector batch, then page by page."""
        imgs, greys = reformat_input_batched(image, n_width, n_height, self.device)
        h_agg, f_agg = self.transform_ua(
            imgs, min_size=min_size, text_threshold=text_threshold, low_text=low_text, link_threshold=link_threshold,
            canvas_size=canvas_size, mag_ratio=mag_ratio, slope_ths=slope_ths, ycenter_ths=ycenter_ths,
            height_ths=height_ths, width_ths=width_ths, add_margin=add_margin, reformat=False, threshold=threshold,
            bbox_min_score=bbox_min_score, bbox_min_size=bbox_min_size, max_candidates=max_candidates)
        return [self.recognize(grey, h_list, f_list, decoder, beamWidth, batch_size, workers, allowlist, blocklist, detail,
                               rotation_info, paragraph, contrast_ths, adjust_contrast, filter_ths, y_ths, x_ths, False,
                               output_format)
                for grey, h_list, f_list in zip(greys, h_agg, f_agg)]

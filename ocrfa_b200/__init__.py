"""ocrfa_b200 - B200-native (sm_100a) recognition hot path of OCR-FA behind the reference's own Python interface.

    from ocrfa_b200 import get_recognizer, get_text, recognizer_predict, Model, CTCLabelConverter

mirrors /root/reference/src/ocrfa/utils.py (get_recognizer :665, get_text :698, recognizer_predict :605,
CTCLabelConverter :1041, get_image_list :1375, get_det_boxes :329, group_text_box :1235), detectors.py (get_textbox
:95), ocr.py (OCR :42) and /root/reference/src/ocrfa/modules/blocks.py (Model :422). The arithmetic is in
libocrfa_b200.so (include/ocrfa_b200.h); there is no CPU fallback.
"""
from . import _lib
from ._lib import OcrfaError
from .charset import (ARABIC_G1_CHARACTERS, ARABIC_G1_SYMBOLS, default_character, default_ignore_char,
                      default_lang_char)
from .converter import CTCLabelConverter
from .craft import CraftNet, get_detector, synth_craft_state_dict
from .crops import calculate_ratio, get_image_list
from .detection import (adjust_result_coordinates, det_components, detector_input, get_det_boxes, get_textbox,
                        group_text_box, test_net)
from .model import Model
from .reader import (OCR, diff, get_paragraph, merge_to_free, reformat_input, reformat_input_batched,
                     set_result_with_confidence)
from .recognition import (AlignCollate, LineBatch, ListDataset, adjust_contrast_grey, contrast_grey, custom_mean,
                          get_recognizer, get_text, recognize_host, recognizer_predict)

__all__ = [
    "Model", "CTCLabelConverter", "get_recognizer", "get_text", "recognizer_predict", "recognize_host",
    "AlignCollate", "ListDataset", "LineBatch", "custom_mean", "contrast_grey", "adjust_contrast_grey",
    "default_character", "default_lang_char", "default_ignore_char", "ARABIC_G1_CHARACTERS", "ARABIC_G1_SYMBOLS",
    "OcrfaError", "get_image_list", "calculate_ratio", "CraftNet", "get_detector",
    "OCR", "reformat_input", "reformat_input_batched", "get_paragraph", "merge_to_free", "set_result_with_confidence",
    "diff", "get_textbox", "test_net", "get_det_boxes", "adjust_result_coordinates", "group_text_box", "detector_input",
    "det_components",
]

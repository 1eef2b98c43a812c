"""Model configurations of the drop-in detector — same keys and values as the reference's
detect_face/utils/config.py: cfg_mnet (:1-16, Retinaface(backbone="mobilenet")) and cfg_re50 (:18-29)."""
from vdesens_b200.arch import CFG_RE50
from vdesens_b200.arch_mnet import CFG_MNET

cfg_mnet = dict(CFG_MNET, loc_weight=2.0, train_image_size=840)
cfg_re50 = dict(CFG_RE50, loc_weight=2.0, train_image_size=840)

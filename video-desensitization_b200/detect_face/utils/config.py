"""Model configuration of the drop-in detector — same keys and values as the reference's
detect_face/utils/config.py:18-29 (cfg_re50).  The MobileNet-0.25 variant (cfg_mnet, config.py:1-16)
is outside the hot path named by BASELINE.json."""
from vdesens_b200.arch import CFG_RE50

cfg_re50 = dict(CFG_RE50, loc_weight=2.0, train_image_size=840)

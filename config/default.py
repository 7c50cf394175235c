"""Drop-in module path of the reference's config/default.py: `from config.default import cfg` yields a
yacs-compatible tree with the same keys and None defaults (reference config/default.py:1-141)."""
from mickey_b200.config import default_cfg

cfg = default_cfg()

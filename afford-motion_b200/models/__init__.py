"""Drop-in `models` package (same import names as /root/reference/models): put `afford-motion_b200/` ahead of the
reference root on PYTHONPATH and the reference's train.py / train_ddp.py / test.py resolve
`models.base.create_model_and_diffusion`, `models.cdm.CDM`, `models.cmdm.CMDM` to the B200-native implementation."""
from models.cdm import *  # noqa: F401,F403  (mirrors reference models/__init__.py:1-2: importing registers the classes)
from models.cmdm import *  # noqa: F401,F403

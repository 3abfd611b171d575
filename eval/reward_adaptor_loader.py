"""Same import path as the reference's scoring API (reference eval/reward_adaptor_loader.py), backed by the B200 engine."""
from llava_reward_b200.reward_adaptor_loader import (inference_process_phi3v, load_reward_adaptor,  # noqa: F401
                                                     preference_compute)

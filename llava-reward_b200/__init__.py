"""llava-reward-b200: B200-native reward-scoring forward path of LLaVA-Reward (Phi-3.5-vision).

Public surface mirrors the reference's scoring API (reference eval/reward_adaptor_loader.py):
`load_reward_adaptor`, `inference_process_phi3v`, `preference_compute`, and a model object
exposing `custom_forward`. All device work runs in hand-written sm_100a CUDA behind the C ABI
declared in `include/llava_reward_b200.h`; there is no CPU fallback.
"""
__version__ = "0.1.0"

"""CPU oracle for the gelslim_depth U-Net hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``gelslim_depth_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and there only as
the checker / the CPU baseline, never as the thing shipped.

Parity status: PINNED.  The reference ships no golden vectors of its own, so
the oracle is pinned against outputs of the reference itself, generated in the build
container by ``tests/golden/make_golden.py`` (which imports ``/root/reference``) and committed
as small fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.
"""
from .unet_oracle import (unet_forward, unet_forward_with_taps, conditioned_state_dict,
                          trainer_init_state_dict, random_init_state_dict, state_dict_digest)
from .processing_oracle import (get_difference_image, area_resample, normalize_tactile_image,
                                denormalize_depth_image, normalize_depth_image,
                                predict_depth_from_RGB, split_fingers, blur_depth_images,
                                preprocess_object_tensors)
from .train_oracle import TrainOracle, mse_loss
from .bf16_sim import loss_and_grads_bf16

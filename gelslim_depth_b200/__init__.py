"""gelslim_depth_b200 -- B200 (sm_100a) implementation of the gelslim_depth U-Net hot path.

Drop-in surface (same names / call signatures as the reference package):
    gelslim_depth_b200.models.unet.UNet
    gelslim_depth_b200.processing_utils.complete_prediction.predict_depth_from_RGB
    gelslim_depth_b200.processing_utils.image_utils / normalization_utils
All compute goes through libgsd_b200.so (include/gsd_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

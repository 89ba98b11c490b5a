"""structuredetector_b200 -- B200-native decoding path for SDNet (laclouis5/StructureDetector).

Drop-in for the reference's ``sdnet.data.Decoder`` (reference: src/sdnet/data/__init__.py:2):

    from structuredetector_b200 import Decoder
    annotations = Decoder(args)(network_outputs)           # list[ImageAnnotation]

The tensor half runs in hand-written sm_100a kernels (``csrc/``) behind a C ABI
(``include/sdnet_decode.h``); there is no CPU fallback.
"""
from .annotations import Box, ImageAnnotation, Keypoint, Object
from .decoders import CoreMLDecoder, CoreMLModel, Decoder, KeypointDecoder, RawDecoder

__all__ = ["Decoder", "CoreMLDecoder", "KeypointDecoder", "RawDecoder", "CoreMLModel", "Keypoint", "Box", "Object", "ImageAnnotation"]
__version__ = "0.1.0"

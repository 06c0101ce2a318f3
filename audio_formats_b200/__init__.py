"""audio_formats_b200 -- B200-native MPEG-1/2/2.5 Layer III granule decode path (drop-in for the
MP3 hot path of AuburnSounds/audio-formats).

The product is the CUDA shared library ``libl3b200.so`` (C-ABI in ``include/l3b200.h``); this
package is a thin ctypes mirror of the reference's AudioStream surface plus the batch entry point.
There is no CPU fallback: every compute call raises when the library or a CUDA device is missing.
"""
from .api import (AudioStream, Context, L3BError, Scan, decode_batch_with_taps, device_count, library_path,  # noqa: F401
                  load_library)
from .pipeline import BatchPipeline  # noqa: F401

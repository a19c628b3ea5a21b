"""B200-native (sm_100a) implementation of hydrus-video-deduplicator's compute hot path:
per-frame PDQ/VPDQ hashing and Hamming similarity search, behind the reference's own Python surface.

    from hydrus_video_deduplicator_b200 import vpdq              # == hvdaccelerators.vpdq
    from hydrus_video_deduplicator_b200.vpdqpy import Vpdq       # == hydrusvideodeduplicator.vpdqpy
    from hydrus_video_deduplicator_b200 import hashing           # == hydrusvideodeduplicator.hashing
    from hydrus_video_deduplicator_b200.search import HashIndex  # brute-force stand-in for db.vptree

All compute runs in libvpdq_b200.so (CUDA, C ABI in include/vpdq_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

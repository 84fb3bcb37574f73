"""TEST INFRASTRUCTURE ONLY — import stub for the un-vendored ``pyloudnorm`` (reference ``environment.yml``).

The reference's ``data/waveform_mixers.py:6`` imports it at module level, but only the retired ``random_loudness_norm``
(``:113-130``, marked "decayed", no caller) uses it; ``SegmentMixer`` does not.  This stub lets the UNMODIFIED reference module be
imported to pin ``oracle/segment_mixer_oracle.py``; touching the loudness meter raises.
"""


class Meter:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("pyloudnorm is not available here (stub under oracle/): BS.1770 loudness is off the path")

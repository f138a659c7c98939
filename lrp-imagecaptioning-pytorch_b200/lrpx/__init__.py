"""lrpx — B200-native LRP kernels (liblrpx.so) and their torch front ends."""
from ._lib import LrpxError, lib  # noqa: F401
from . import ops  # noqa: F401

"""Drop-in mirror of the reference's LRPtools package (same names, argument meaning and error
behaviour) whose arithmetic runs in liblrpx.so (hand-written sm_100a CUDA)."""

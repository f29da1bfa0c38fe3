"""Host-side mirror of the reference's pkg/gpu/nvidia (same names, argument meaning and error
behaviour), over the C ABI of libgpushare_b200.so."""

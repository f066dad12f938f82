"""Bindings to libnbdt_b200.so (ctypes over the C ABI in include/nbdt_b200.h)."""

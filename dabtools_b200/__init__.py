"""dabtools_b200 -- B200-native (sm_100a) implementation of the dabtools receive hot path.

The product is the C-ABI shared library ``libdabgpu.so`` (sources in ``csrc/``, headers in
``include/``).  The Python modules here are a thin ctypes mirror of that ABI (``lib``), the table
accessors (``tables``), the host-side mirror of the reference's call interface (``refapi``) and the
synthetic Mode I transmitter used to manufacture inputs (``synth``).
"""
__all__ = ["lib", "tables", "synth"]

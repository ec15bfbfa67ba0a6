"""compressai.ans is a C++ rANS coder used only by the serial test() path
(LiftingBasedDWT_net.py:458-556), which is out of scope; importing must work."""


class BufferedRansEncoder:
    def __init__(self):
        raise NotImplementedError("compressai.ans (C++ rANS) is not available offline")


class RansDecoder:
    def __init__(self):
        raise NotImplementedError("compressai.ans (C++ rANS) is not available offline")

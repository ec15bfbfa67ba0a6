from oracle.thirdparty import rgb2ycbcr, ycbcr2rgb


class RGB2YCbCr:
    def __call__(self, rgb):
        return rgb2ycbcr(rgb)


class YCbCr2RGB:
    def __call__(self, ycbcr):
        return ycbcr2rgb(ycbcr)

// Convolution geometry shared by the FFMA and tcgen05 paths.
#pragma once

struct ConvGeom {
    int N, H, W, C;      // input  (channels-last)
    int OH, OW, OC;      // output (channels-last)
    int KH, KW, sh, sw, pad;
};

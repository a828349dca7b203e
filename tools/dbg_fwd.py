import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiccup_b200 import compression, _lib
from oracle import hiccup_oracle as orc
for shape in ((32,32),(64,64),(426,640),(40,24)):
    rgb = orc.synthetic_image(shape[0], shape[1], 1)
    try:
        out = compression.jpeg_compression(rgb)
        want = orc.jpeg_compression(rgb)
        print(shape, 'ok', [int((out.as_dict[c]!=want[c]).sum()) for c in orc.CHANNELS], compression.LAST_STATS)
    except Exception as e:
        print(shape, 'ERR', e)
        break

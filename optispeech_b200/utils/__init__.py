from .model import denormalize, make_non_pad_mask, make_pad_mask, normalize, pad_list, safe_log, sequence_mask
from .segments import get_random_segments, get_segments, get_segments_numpy

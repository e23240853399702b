"""Record-range filter used by data-position resume (reference: python/edl/utils/data_filter.py)."""


def is_processed(record_no, processed_ranges):
    """``processed_ranges``: iterable of (begin, end) inclusive ranges already consumed."""
    for b, e in processed_ranges or ():
        if b <= record_no <= e:
            return True
    return False


def merge_ranges(ranges):
    """Coalesce overlapping / adjacent inclusive ranges."""
    out = []
    for b, e in sorted(ranges):
        if out and b <= out[-1][1] + 1:
            out[-1][1] = max(out[-1][1], e)
        else:
            out.append([b, e])
    return [tuple(x) for x in out]

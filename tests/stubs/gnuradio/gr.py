"""TEST STAND-IN for gnuradio.gr: the Python block API surface gr_adapter.py uses (GNU Radio 3.8 gateway blocks):
sync_block / basic_block with nitems_read / nitems_written / consume / consume_each / add_item_tag /
get_tags_in_range / set_output_multiple, and a tag record. The test plays scheduler."""


class tag_t:
    def __init__(self, offset=0, key=None, value=None, srcid=None):
        self.offset, self.key, self.value, self.srcid = offset, key, value, srcid


class _block:
    def __init__(self, name, in_sig, out_sig):
        self._name, self.in_sig, self.out_sig = name, in_sig or [], out_sig or []
        self._nread = [0] * len(self.in_sig)
        self._nwritten = [0] * max(1, len(self.out_sig))
        self._consumed = [0] * len(self.in_sig)
        self._in_tags = [[] for _ in self.in_sig]          # the scheduler stand-in fills these
        self.out_tags = []
        self.output_multiple = 1

    def name(self): return self._name
    def nitems_read(self, port): return self._nread[port]
    def nitems_written(self, port): return self._nwritten[port]
    def set_output_multiple(self, m): self.output_multiple = int(m)
    def consume(self, port, n): self._consumed[port] += int(n)

    def consume_each(self, n):
        for p in range(len(self._consumed)):
            self._consumed[p] += int(n)

    def add_item_tag(self, port, offset, key, value, srcid=None):
        self.out_tags.append(tag_t(int(offset), key, value, srcid))

    def get_tags_in_range(self, port, start, end, key=None):
        return [t for t in self._in_tags[port] if start <= t.offset < end and (key is None or t.key == key)]

    def start(self): return True
    def stop(self): return True


class sync_block(_block):
    pass


class basic_block(_block):
    pass


def message_from_string(s):
    class _M:
        def to_string(self): return s
    return _M()

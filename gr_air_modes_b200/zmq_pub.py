"""ZeroMQ "dl_data" publisher: a queue object for rx_path that puts every slicer message
on a ZMQ PUB socket exactly as modes_radio does (python/radio.py:79-87: zmq_pubsub_iface["dl_data"] =
msg.to_string(), python/zmq_socket.py:82-101: send_multipart([key, value])). Existing subscribers
(modes_rx -a/--remote, modes_gui; apps/modes_rx:58-63) can connect unchanged.

    q = zmq_queue(["tcp://*:5556"])          # instead of gr.msg_queue()
    rx = rx_path(4e6, 7.0, q, use_pmf=True)
"""
from __future__ import annotations


class zmq_queue:
    """Has the msg_queue methods rx_path/slicer use (handle, insert_tail); publishes instead of queueing."""

    def __init__(self, pubaddr, context=None, key: str = "dl_data"):
        import zmq
        self._ctx = context or zmq.Context.instance()
        self._sock = self._ctx.socket(zmq.PUB)
        for addr in ([pubaddr] if isinstance(pubaddr, str) else pubaddr):
            self._sock.bind(addr)
        self._key = key.encode("ascii")
        self.sent = 0

    def handle(self, msg) -> None:
        text = msg.to_string()
        self._sock.send_multipart([self._key, text if isinstance(text, bytes) else text.encode("ascii")])
        self.sent += 1

    insert_tail = handle

    def empty_p(self) -> bool:
        return True

    def count(self) -> int:
        return 0

    def close(self) -> None:
        self._sock.close(linger=0)

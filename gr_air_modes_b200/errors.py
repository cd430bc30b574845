"""The exceptions the reference's decode functions raise (python/exceptions.py), for code that catches them around
`cpr_decoder.decode()`. The batch API reports the same conditions as bits of amb_fields.status instead."""


class ADSBError(Exception):
    """Base of everything the reference's parser / CPR / altitude code raises."""


class MetricAltError(ADSBError):
    """decode_alt: M bit set (altitude.py:32-43)."""


class NoHandlerError(ADSBError):
    """No field table for this DF / FTC / BDS register (parse.py:52-68)."""

    def __init__(self, msgtype=None):
        super().__init__(msgtype)
        self.msgtype = msgtype


class CPRNoPositionError(ADSBError):
    """No live even/odd pair yet, or a surface report without a receiver location (cpr.py:97-99, 231)."""


class CPRBoundaryStraddleError(CPRNoPositionError):
    """The even and odd reports lie in different longitude zones (cpr.py:120-121)."""

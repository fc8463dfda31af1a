"""Host-side mirror of the reference's modules/ interface for the hot path (same names, argument
meaning and return values as ksahlin/NGSpeciesID v0.3.1), backed by libngsid.so on a B200."""

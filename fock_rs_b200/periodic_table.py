"""Element table for the host mirror (reference: basis_set/src/periodic_table.rs).

`AtomType` in the reference is an enum whose discriminant is Z (periodic_table.rs:8-263); only the
eight symbols below parse (`from_symbol`, :284-296) -- anything else panics ("not implemented yet").
"""

_SYMBOL_TO_Z = {"H": 1, "He": 2, "C": 6, "N": 7, "O": 8, "F": 9, "Fe": 26, "B": 5}


def from_symbol(symbol: str) -> int:
    try:
        return _SYMBOL_TO_Z[symbol]
    except KeyError:
        raise NotImplementedError(f"not implemented yet: element symbol {symbol!r}") from None

"""`Data.Text.CaseSensitivity` (src/Data/Text/CaseSensitivity.hs:14-16)."""
import enum


class CaseSensitivity(enum.IntEnum):
    CaseSensitive = 0
    IgnoreCase = 1


CaseSensitive = CaseSensitivity.CaseSensitive
IgnoreCase = CaseSensitivity.IgnoreCase

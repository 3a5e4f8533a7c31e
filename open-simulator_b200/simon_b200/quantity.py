"""resource.Quantity semantics needed by the placement path (host side).

Restates the *vendored* apimachinery code of the reference (the vendored tree, not upstream, is the
specification — SURVEY.md §8c):
  ParseQuantity            vendor/k8s.io/apimachinery/pkg/api/resource/quantity.go:247-372
  Quantity.Add/Sub/Cmp     quantity.go:560-600, amount.go:108-204
  Value / MilliValue       quantity.go (ScaledValue), amount.go:91-99, math.go:181-213
  AsApproximateFloat64     quantity.go:449-474   (including its base-2 scaling of decimal exponents
                                                  for BinarySI quantities held as inf.Dec — kept as is)
Go float64 arithmetic == Python float arithmetic (IEEE-754 binary64, round-to-nearest-even).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

DECIMAL_EXPONENT = "DecimalExponent"
BINARY_SI = "BinarySI"
DECIMAL_SI = "DecimalSI"

_MAX_INT64 = (1 << 63) - 1
_MIN_INT64 = -(1 << 63)
_MAX_INT64_FACTORS = 18
_NANO = -9

_BINARY_SUFFIX = {"Ki": 10, "Mi": 20, "Gi": 30, "Ti": 40, "Pi": 50, "Ei": 60}
_DECIMAL_SUFFIX = {"n": -9, "u": -6, "m": -3, "": 0, "k": 3, "M": 6, "G": 9, "T": 12, "P": 15, "E": 18}


class QuantityError(ValueError):
    pass


def _fits(v: int) -> bool:
    return _MIN_INT64 <= v <= _MAX_INT64


def _interpret_suffix(suf: str):
    """quantitySuffixer.interpret (suffix.go): returns (base, exponent, format)."""
    if suf in _DECIMAL_SUFFIX:
        return 10, _DECIMAL_SUFFIX[suf], DECIMAL_SI
    if suf in _BINARY_SUFFIX:
        return 2, _BINARY_SUFFIX[suf], BINARY_SI
    if len(suf) > 1 and suf[0] in "eE":
        try:
            return 10, int(suf[1:]), DECIMAL_EXPONENT
        except ValueError:
            pass
    raise QuantityError("unable to parse quantity's suffix")


def _parse_quantity_string(s: str):
    """parseQuantityString (quantity.go:148-245)."""
    positive = True
    pos = 0
    end = len(s)
    if pos < end and s[0] in "+-":
        positive = s[0] != "-"
        pos += 1
    # strip leading zeros
    i = pos
    while True:
        if i >= end:
            return positive, "0", "0", "", ""
        if s[i] == "0":
            pos += 1
            i += 1
        else:
            break
    # numerator
    i = pos
    num = None
    while True:
        if i >= end:
            return positive, s[0:end], s[pos:end], "", ""
        if s[i].isdigit() and s[i].isascii():
            i += 1
        else:
            num = s[pos:i]
            pos = i
            break
    if len(num) == 0:
        num = "0"
    denom = ""
    if pos < end and s[pos] == ".":
        pos += 1
        i = pos
        while True:
            if i >= end:
                return positive, s[0:end], num, s[pos:end], ""
            if s[i].isdigit() and s[i].isascii():
                i += 1
            else:
                denom = s[pos:i]
                pos = i
                break
    value = s[0:pos]
    suffix_start = pos
    i = pos
    while True:
        if i >= end:
            return positive, value, num, denom, s[suffix_start:end]
        if s[i] not in "eEinumkKMGTP":
            pos = i
            break
        i += 1
    if pos < end and s[pos] in "+-":
        pos += 1
    i = pos
    while True:
        if i >= end:
            return positive, value, num, denom, s[suffix_start:end]
        if s[i].isdigit() and s[i].isascii():
            i += 1
        else:
            break
    raise QuantityError("quantities must match the regular expression")


_PARSE_CACHE: dict = {}      # quantity strings repeat heavily across nodes and pod templates


@dataclass
class Quantity:
    """Either an int64Amount (value * 10^scale) or an inf.Dec (unscaled * 10^-dscale)."""
    value: int = 0          # int64Amount.value
    scale: int = 0          # int64Amount.scale
    dec_unscaled: Optional[int] = None  # inf.Dec unscaled big int (None => int64Amount form)
    dec_scale: int = 0      # inf.Dec scale (digits after the point)
    format: str = ""

    # -- constructors -------------------------------------------------------------------------
    @staticmethod
    def parse(s) -> "Quantity":
        if isinstance(s, Quantity):
            return s.copy()
        if isinstance(s, bool):
            raise QuantityError("bool is not a quantity")
        if isinstance(s, int):
            s = str(s)
        elif isinstance(s, float):
            s = repr(s) if s != int(s) else str(int(s))
        s = s.strip()
        if len(s) == 0:
            raise QuantityError("empty quantity")
        if s == "0":
            return Quantity(format=DECIMAL_SI)
        hit = _PARSE_CACHE.get(s)
        if hit is not None:
            return hit.copy()
        q = Quantity._parse_uncached(s)
        if len(_PARSE_CACHE) < 65536:
            _PARSE_CACHE[s] = q.copy()
        return q

    @staticmethod
    def _parse_uncached(s: str) -> "Quantity":
        positive, value, num, denom, suf = _parse_quantity_string(s)
        base, exponent, fmt = _interpret_suffix(suf)
        precision = 0
        scale = 0
        mantissa = 1
        if fmt in (DECIMAL_EXPONENT, DECIMAL_SI):
            scale = exponent
            precision = _MAX_INT64_FACTORS - (len(num) + len(denom))
        else:
            scale = 0
            if exponent >= 0 and len(denom) == 0:
                mantissa = 1 << exponent
                # float32(exponent)*3/10 truncated toward zero
                precision = 15 - len(num) - int(float(exponent) * 3 / 10) - 1
            else:
                precision = -1
        if precision >= 0:
            scale -= len(denom)
            if scale >= _NANO:
                shifted = num + denom
                v = int(shifted)
                if _fits(v):
                    result = v * mantissa
                    if _fits(result) and (v == 0 or result // mantissa == v):
                        if not positive:
                            result = -result
                        return Quantity(value=result, scale=scale, format=fmt)
        # inf.Dec path
        digits = value.lstrip("+-")
        if "." in digits:
            ip, fp = digits.split(".", 1)
        else:
            ip, fp = digits, ""
        unscaled = int((ip + fp) or "0")
        dscale = len(fp)
        if base == 10:
            dscale = dscale - exponent  # amount.SetScale(amount.Scale() + Scale(exponent).infScale()) ; infScale = -exponent
        else:
            unscaled = unscaled * (1 << exponent)
        # (sign handled below; amount is non-negative here)
        if unscaled != 0:
            # amount.Round(amount, Nano.infScale() == 9, RoundUp): quantise to exactly 9 fractional digits
            unscaled, dscale = _round_up_to_scale(unscaled, dscale, 9)
        if fmt == BINARY_SI:
            max_unscaled, max_scale = _MAX_INT64, 0
            if _dec_cmp(unscaled, dscale, max_unscaled, max_scale) > 0:
                unscaled, dscale = max_unscaled, max_scale
            if _dec_cmp(unscaled, dscale, 1, 0) < 0 and unscaled > 0:
                fmt = DECIMAL_SI
        if not positive:
            unscaled = -unscaled
        return Quantity(dec_unscaled=unscaled, dec_scale=dscale, format=fmt)

    @staticmethod
    def new(value: int, fmt: str) -> "Quantity":
        return Quantity(value=value, scale=0, format=fmt)

    @staticmethod
    def new_milli(value: int, fmt: str) -> "Quantity":
        return Quantity(value=value, scale=-3, format=fmt)

    def copy(self) -> "Quantity":
        return Quantity(self.value, self.scale, self.dec_unscaled, self.dec_scale, self.format)

    # -- helpers ------------------------------------------------------------------------------
    def is_dec(self) -> bool:
        return self.dec_unscaled is not None

    def _as_dec(self):
        """AsDec: (unscaled, scale) with value == unscaled * 10^-scale."""
        if self.is_dec():
            return self.dec_unscaled, self.dec_scale
        return self.value, -self.scale

    def is_zero(self) -> bool:
        if self.is_dec():
            return self.dec_unscaled == 0
        return self.value == 0

    def sign(self) -> int:
        v = self.dec_unscaled if self.is_dec() else self.value
        return (v > 0) - (v < 0)

    # -- arithmetic ---------------------------------------------------------------------------
    def _i_add(self, bvalue: int, bscale: int) -> bool:
        """int64Amount.Add (amount.go:162-199); mutates self, returns False on overflow."""
        if bvalue == 0:
            return True
        if self.value == 0:
            self.value, self.scale = bvalue, bscale
            return True
        if self.scale == bscale:
            c = self.value + bvalue
            if not _fits(c):
                return False
            self.value = c
            return True
        if self.scale > bscale:
            c = self.value * 10 ** (self.scale - bscale)
            if not _fits(c):
                return False
            c += bvalue
            if not _fits(c):
                return False
            self.scale = bscale
            self.value = c
            return True
        c = bvalue * 10 ** (bscale - self.scale)
        if not _fits(c):
            return False
        c = self.value + c
        if not _fits(c):
            return False
        self.value = c
        return True

    def add(self, y: "Quantity") -> None:
        """Quantity.Add (quantity.go:560-575)."""
        if self.is_zero():
            self.format = y.format
        if not self.is_dec() and not y.is_dec():
            saved = (self.value, self.scale)
            if self._i_add(y.value, y.scale):
                return
            self.value, self.scale = saved
        au, asc = self._as_dec()
        bu, bsc = y._as_dec()
        self.dec_unscaled, self.dec_scale = _dec_add(au, asc, bu, bsc)
        self.value, self.scale = 0, 0

    def sub(self, y: "Quantity") -> None:
        """Quantity.Sub (quantity.go:579-588)."""
        if self.is_zero():
            self.format = y.format
        if not self.is_dec() and not y.is_dec():
            saved = (self.value, self.scale)
            if self._i_add(-y.value, y.scale):
                return
            self.value, self.scale = saved
        au, asc = self._as_dec()
        bu, bsc = y._as_dec()
        self.dec_unscaled, self.dec_scale = _dec_add(au, asc, -bu, bsc)
        self.value, self.scale = 0, 0

    def cmp(self, y: "Quantity") -> int:
        au, asc = self._as_dec()
        bu, bsc = y._as_dec()
        return _dec_cmp(au, asc, bu, bsc)

    # -- conversions --------------------------------------------------------------------------
    def scaled_value(self, scale: int) -> int:
        """ScaledValue: value / 10^scale rounded up (away from zero) to an int64."""
        if self.is_dec():
            u, dsc = self.dec_unscaled, self.dec_scale
            # want u * 10^-dsc / 10^scale
            exp = -dsc - scale
            if exp >= 0:
                return u * 10 ** exp
            d = 10 ** (-exp)
            q, r = divmod(abs(u), d)
            if r:
                q += 1
            return q if u >= 0 else -q
        if self.scale < scale:
            return _negative_scale(self.value, scale - self.scale)
        return self.value * 10 ** (self.scale - scale)

    def int_value(self) -> int:
        return self.scaled_value(0)

    def milli_value(self) -> int:
        return self.scaled_value(-3)

    def as_approximate_float64(self) -> float:
        """Quantity.AsApproximateFloat64 (quantity.go:449-474)."""
        if self.is_dec():
            base = _bigint_to_float(self.dec_unscaled)
            exponent = -self.dec_scale
        else:
            base = float(self.value)
            exponent = self.scale
        if exponent == 0:
            return base
        if self.format in (DECIMAL_EXPONENT, DECIMAL_SI):
            return base * go_pow10(exponent)
        if 0 < exponent < 7:
            return base * float(1 << (exponent * 10))
        return base * math.ldexp(1.0, exponent * 10)

    def key(self):
        """Hashable identity (what AsApproximateFloat64 / Sub can observe)."""
        return (self.value, self.scale, self.dec_unscaled, self.dec_scale, self.format)


def go_pow10(n: int) -> float:
    """Go math.Pow10: table product/quotient (src/math/pow10.go)."""
    if 0 <= n <= 308:
        return float(10 ** ((n // 32) * 32)) * float(10 ** (n % 32))
    if -323 <= n <= 0:
        m = -n
        return _neg32(m // 32) / float(10 ** (m % 32))
    return math.inf if n > 0 else 0.0


def _neg32(i: int) -> float:
    # pow10negtab32[i] = 1e-(32 i) as a correctly rounded literal
    from fractions import Fraction
    if i == 0:
        return 1.0
    return float(Fraction(1, 10 ** (32 * i)))


def _bigint_to_float(v: int) -> float:
    # big.Float(prec=len bits).SetInt(v).Float64() rounds to nearest even == Python int->float
    return float(v)


def _negative_scale(base: int, scale: int) -> int:
    """negativeScaleInt64 (math.go:181-213): divide by 10^scale rounding away from zero."""
    if scale == 0:
        return base
    value = base
    fraction = False
    for _ in range(scale):
        if not fraction and abs(value) % 10 != 0:   # Go % truncates toward zero
            fraction = True
        value = abs(value) // 10 if value >= 0 else -(abs(value) // 10)   # Go / truncates toward zero
        if value == 0:
            if fraction:
                return 1 if base > 0 else -1
            return 0
    if fraction:
        value = value + 1 if base > 0 else value - 1
    return value


def _round_up_to_scale(unscaled: int, dscale: int, target: int):
    if dscale <= target:
        return unscaled * 10 ** (target - dscale), target
    d = 10 ** (dscale - target)
    q, r = divmod(abs(unscaled), d)
    if r:
        q += 1
    return (q if unscaled >= 0 else -q), target


def _dec_add(au, asc, bu, bsc):
    s = max(asc, bsc)
    return au * 10 ** (s - asc) + bu * 10 ** (s - bsc), s


def _dec_cmp(au, asc, bu, bsc) -> int:
    s = max(asc, bsc)
    a = au * 10 ** (s - asc)
    b = bu * 10 ** (s - bsc)
    return (a > b) - (a < b)

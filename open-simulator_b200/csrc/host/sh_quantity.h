// resource.Quantity as far as the placement path observes it (native host side; mirrors simon_b200/quantity.py).
// Restates the reference's VENDORED apimachinery (paths relative to the reference tree):
//   ParseQuantity            vendor/k8s.io/apimachinery/pkg/api/resource/quantity.go:247-372
//   parseQuantityString      quantity.go:148-245 ; quantitySuffixer.interpret suffix.go
//   Quantity.Add/Sub/Cmp     quantity.go:560-600, amount.go:108-204
//   Value / MilliValue       quantity.go (ScaledValue), amount.go:91-99, math.go:181-213
//   AsApproximateFloat64     quantity.go:449-474
// inf.Dec amounts are held as a 128-bit unscaled integer; a quantity that needs more (≈ 38 digits) is refused, not rounded.
#pragma once
#include <cmath>
#include <string>
#include <unordered_map>

#include "sh_json.h"

namespace sh {

typedef __int128 i128;
static const long long kMaxI64 = 9223372036854775807LL;
static const long long kMinI64 = (-9223372036854775807LL - 1);

enum QFormat : uint8_t { QF_NONE = 0, QF_DECIMAL_EXPONENT, QF_BINARY_SI, QF_DECIMAL_SI };

struct QuantityError : Error { explicit QuantityError(const std::string &m) : Error("quantity: " + m) {} };

inline bool fits64(i128 v) { return v >= (i128)kMinI64 && v <= (i128)kMaxI64; }

inline i128 mul_checked(i128 a, i128 b) {
    i128 r;
    if (__builtin_mul_overflow(a, b, &r)) throw QuantityError("value exceeds 128 bits");
    return r;
}
inline i128 add_checked(i128 a, i128 b) {
    i128 r;
    if (__builtin_add_overflow(a, b, &r)) throw QuantityError("value exceeds 128 bits");
    return r;
}
inline i128 pow10_128(int n) {
    if (n < 0 || n > 38) throw QuantityError("scale out of range");
    i128 r = 1;
    for (int i = 0; i < n; i++) r = mul_checked(r, 10);
    return r;
}
inline i128 abs128(i128 v) { return v < 0 ? -v : v; }

// Go math.Pow10 (src/math/pow10.go): pow10postab32[n/32] * pow10tab[n%32] resp. pow10negtab32[-n/32] / pow10tab[-n%32]; the table
// entries are correctly rounded decimal literals, which is what strtod yields
inline double lit_pow10(int n) {
    char buf[16];
    snprintf(buf, sizeof buf, "1e%d", n);
    return strtod(buf, nullptr);
}
inline double go_pow10(int n) {
    if (0 <= n && n <= 308) return lit_pow10((n / 32) * 32) * lit_pow10(n % 32);
    if (-323 <= n && n <= 0) { int m = -n; return lit_pow10(-(m / 32) * 32) / lit_pow10(m % 32); }
    return n > 0 ? INFINITY : 0.0;
}

struct Quantity {
    long long value = 0;     // int64Amount.value
    int scale = 0;           // int64Amount.scale
    bool dec = false;        // inf.Dec form: dec_unscaled * 10^-dec_scale
    i128 dec_unscaled = 0;
    int dec_scale = 0;
    QFormat format = QF_NONE;

    bool is_zero() const { return dec ? dec_unscaled == 0 : value == 0; }
    void as_dec(i128 &u, int &sc) const {
        if (dec) { u = dec_unscaled; sc = dec_scale; }
        else { u = value; sc = -scale; }
    }
    // hashable identity (what AsApproximateFloat64 / Sub can observe)
    std::string key() const {
        char buf[128];
        if (!dec) snprintf(buf, sizeof buf, "i%lld,%d,%d", value, scale, (int)format);
        else snprintf(buf, sizeof buf, "d%lld:%llu,%d,%d", (long long)(dec_unscaled >> 64), (unsigned long long)dec_unscaled, dec_scale, (int)format);
        return buf;
    }

    // int64Amount.Add (amount.go:162-199): false on overflow, *this untouched then
    bool i_add(long long bvalue, int bscale) {
        if (bvalue == 0) return true;
        if (value == 0) { value = bvalue; scale = bscale; return true; }
        if (scale == bscale) {
            i128 c = (i128)value + bvalue;
            if (!fits64(c)) return false;
            value = (long long)c;
            return true;
        }
        if (scale > bscale) {
            int d = scale - bscale;
            if (d > 18) return false;
            i128 c = (i128)value * pow10_128(d);
            if (!fits64(c)) return false;
            c += bvalue;
            if (!fits64(c)) return false;
            scale = bscale;
            value = (long long)c;
            return true;
        }
        int d = bscale - scale;
        if (d > 18) return false;
        i128 c = (i128)bvalue * pow10_128(d);
        if (!fits64(c)) return false;
        c = (i128)value + c;
        if (!fits64(c)) return false;
        value = (long long)c;
        return true;
    }
    static void dec_add(i128 au, int asc, i128 bu, int bsc, i128 &ru, int &rsc) {
        int s = asc > bsc ? asc : bsc;
        ru = add_checked(mul_checked(au, pow10_128(s - asc)), mul_checked(bu, pow10_128(s - bsc)));
        rsc = s;
    }
    static int dec_cmp(i128 au, int asc, i128 bu, int bsc) {
        int s = asc > bsc ? asc : bsc;
        i128 a = mul_checked(au, pow10_128(s - asc)), b = mul_checked(bu, pow10_128(s - bsc));
        return (a > b) - (a < b);
    }
    // Quantity.Add / Sub (quantity.go:560-588)
    void add_signed(const Quantity &y, int sign) {
        if (is_zero()) format = y.format;
        if (!dec && !y.dec) {
            if (!(sign < 0 && y.value == kMinI64)) {
                long long sv = value;
                int ss = scale;
                if (i_add(sign < 0 ? -y.value : y.value, y.scale)) return;
                value = sv;
                scale = ss;
            }
        }
        i128 au, bu;
        int asc, bsc;
        as_dec(au, asc);
        y.as_dec(bu, bsc);
        dec_add(au, asc, sign < 0 ? -bu : bu, bsc, dec_unscaled, dec_scale);
        dec = true;
        value = 0;
        scale = 0;
    }
    void add(const Quantity &y) { add_signed(y, 1); }
    void sub(const Quantity &y) { add_signed(y, -1); }
    int cmp(const Quantity &y) const {
        i128 au, bu;
        int asc, bsc;
        as_dec(au, asc);
        y.as_dec(bu, bsc);
        return dec_cmp(au, asc, bu, bsc);
    }

    // negativeScaleInt64 (math.go:181-213): divide by 10^scale rounding away from zero
    static long long negative_scale(long long base, int sc) {
        if (sc == 0) return base;
        long long v = base;
        bool fraction = false;
        for (int i = 0; i < sc; i++) {
            if (!fraction && v % 10 != 0) fraction = true;
            v = v / 10;
            if (v == 0) {
                if (fraction) return base > 0 ? 1 : -1;
                return 0;
            }
        }
        if (fraction) v = base > 0 ? v + 1 : v - 1;
        return v;
    }
    // ScaledValue: value / 10^scale rounded up (away from zero); results beyond int64 are refused
    long long scaled_value(int sc) const {
        i128 r;
        if (dec) {
            int exp = -dec_scale - sc;
            if (exp >= 0) r = mul_checked(dec_unscaled, pow10_128(exp));
            else {
                if (-exp > 38) r = dec_unscaled == 0 ? 0 : (dec_unscaled > 0 ? 1 : -1);
                else {
                    i128 d = pow10_128(-exp), a = abs128(dec_unscaled), q = a / d;
                    if (a % d) q += 1;
                    r = dec_unscaled >= 0 ? q : -q;
                }
            }
        } else if (scale < sc) {
            return negative_scale(value, sc - scale);
        } else {
            r = mul_checked(value, pow10_128(scale - sc));
        }
        if (!fits64(r)) throw QuantityError("value does not fit int64");
        return (long long)r;
    }
    long long int_value() const { return scaled_value(0); }
    long long milli_value() const { return scaled_value(-3); }

    // Quantity.AsApproximateFloat64 (quantity.go:449-474), including its base-2 scaling of decimal exponents for BinarySI
    double as_approximate_float64() const {
        double base;
        int exponent;
        if (dec) { base = (double)dec_unscaled; exponent = -dec_scale; }     // big.Float -> float64: round to nearest even
        else { base = (double)value; exponent = scale; }
        if (exponent == 0) return base;
        if (format == QF_DECIMAL_EXPONENT || format == QF_DECIMAL_SI) return base * go_pow10(exponent);
        if (0 < exponent && exponent < 7) return base * (double)(1ULL << (exponent * 10));
        return base * std::ldexp(1.0, exponent * 10);
    }
};

inline bool ascii_digit(char c) { return c >= '0' && c <= '9'; }

// parseQuantityString (quantity.go:148-245)
inline void parse_quantity_string(const std::string &s, bool &positive, std::string &value, std::string &num, std::string &denom,
                                  std::string &suffix) {
    positive = true;
    size_t pos = 0, end = s.size();
    value.clear(); num.clear(); denom.clear(); suffix.clear();
    if (pos < end && (s[0] == '+' || s[0] == '-')) { positive = s[0] != '-'; pos++; }
    size_t i = pos;
    while (true) {      // strip leading zeros
        if (i >= end) { value = "0"; num = "0"; return; }
        if (s[i] == '0') { pos++; i++; } else break;
    }
    i = pos;
    while (true) {      // numerator
        if (i >= end) { value = s.substr(0, end); num = s.substr(pos, end - pos); return; }
        if (ascii_digit(s[i])) i++;
        else { num = s.substr(pos, i - pos); pos = i; break; }
    }
    if (num.empty()) num = "0";
    if (pos < end && s[pos] == '.') {
        pos++;
        i = pos;
        while (true) {
            if (i >= end) { value = s.substr(0, end); denom = s.substr(pos, end - pos); return; }
            if (ascii_digit(s[i])) i++;
            else { denom = s.substr(pos, i - pos); pos = i; break; }
        }
    }
    value = s.substr(0, pos);
    size_t suffix_start = pos;
    i = pos;
    while (true) {
        if (i >= end) { suffix = s.substr(suffix_start); return; }
        if (!strchr("eEinumkKMGTP", s[i])) { pos = i; break; }
        i++;
    }
    if (pos < end && (s[pos] == '+' || s[pos] == '-')) pos++;
    i = pos;
    while (true) {
        if (i >= end) { suffix = s.substr(suffix_start); return; }
        if (ascii_digit(s[i])) i++;
        else break;
    }
    throw QuantityError("quantities must match the regular expression: '" + s + "'");
}

inline void interpret_suffix(const std::string &suf, int &base, int &exponent, QFormat &fmt) {
    static const struct { const char *s; int e; } dec[] = {{"n", -9}, {"u", -6}, {"m", -3}, {"", 0}, {"k", 3}, {"M", 6}, {"G", 9}, {"T", 12},
                                                           {"P", 15}, {"E", 18}};
    static const struct { const char *s; int e; } bin[] = {{"Ki", 10}, {"Mi", 20}, {"Gi", 30}, {"Ti", 40}, {"Pi", 50}, {"Ei", 60}};
    for (auto &d : dec) if (suf == d.s) { base = 10; exponent = d.e; fmt = QF_DECIMAL_SI; return; }
    for (auto &b : bin) if (suf == b.s) { base = 2; exponent = b.e; fmt = QF_BINARY_SI; return; }
    if (suf.size() > 1 && (suf[0] == 'e' || suf[0] == 'E')) {
        // Python int(): optional sign, digits (underscores / spaces are not produced by parseQuantityString's character classes)
        size_t k = 1;
        if (suf[k] == '+' || suf[k] == '-') k++;
        bool ok = k < suf.size();
        for (size_t q = k; q < suf.size(); q++) ok = ok && ascii_digit(suf[q]);
        if (ok && suf.size() - k <= 6) { base = 10; exponent = atoi(suf.c_str() + 1); fmt = QF_DECIMAL_EXPONENT; return; }
    }
    throw QuantityError("unable to parse quantity's suffix '" + suf + "'");
}

inline i128 digits128(const std::string &d) {
    i128 v = 0;
    for (char c : d) v = add_checked(mul_checked(v, 10), c - '0');
    return v;
}

inline Quantity parse_quantity_uncached(const std::string &s) {
    bool positive;
    std::string value, num, denom, suf;
    parse_quantity_string(s, positive, value, num, denom, suf);
    int base, exponent;
    QFormat fmt;
    interpret_suffix(suf, base, exponent, fmt);
    int precision, scale = 0;
    i128 mantissa = 1;
    if (fmt == QF_DECIMAL_EXPONENT || fmt == QF_DECIMAL_SI) {
        scale = exponent;
        precision = 18 - (int)(num.size() + denom.size());
    } else {
        scale = 0;
        if (exponent >= 0 && denom.empty()) {
            mantissa = (i128)1 << exponent;
            precision = 15 - (int)num.size() - (exponent * 3 / 10) - 1;      // int(float32(exponent)*3/10): exact for 10..60
        } else precision = -1;
    }
    if (precision >= 0) {
        scale -= (int)denom.size();
        if (scale >= -9) {
            i128 v = digits128(num + denom);       // <= 18 digits here
            if (fits64(v)) {
                i128 result = v * mantissa;
                if (fits64(result)) {
                    Quantity q;
                    q.value = (long long)(positive ? result : -result);
                    q.scale = scale;
                    q.format = fmt;
                    return q;
                }
            }
        }
    }
    // inf.Dec path
    std::string digits = value;
    while (!digits.empty() && (digits[0] == '+' || digits[0] == '-')) digits.erase(0, 1);
    std::string ip = digits, fp;
    size_t dot = digits.find('.');
    if (dot != std::string::npos) { ip = digits.substr(0, dot); fp = digits.substr(dot + 1); }
    i128 unscaled = digits128(ip + fp);
    int dscale = (int)fp.size();
    if (base == 10) dscale -= exponent;
    else unscaled = mul_checked(unscaled, (i128)1 << exponent);
    if (unscaled != 0) {
        // amount.Round(amount, 9, RoundUp): quantise to exactly 9 fractional digits
        if (dscale <= 9) unscaled = mul_checked(unscaled, pow10_128(9 - dscale));
        else {
            if (dscale - 9 > 38) unscaled = 1;
            else {
                i128 d = pow10_128(dscale - 9), q = unscaled / d;
                if (unscaled % d) q += 1;
                unscaled = q;
            }
        }
        dscale = 9;
    }
    if (fmt == QF_BINARY_SI) {
        if (Quantity::dec_cmp(unscaled, dscale, kMaxI64, 0) > 0) { unscaled = kMaxI64; dscale = 0; }
        if (Quantity::dec_cmp(unscaled, dscale, 1, 0) < 0 && unscaled > 0) fmt = QF_DECIMAL_SI;
    }
    Quantity q;
    q.dec = true;
    q.dec_unscaled = positive ? unscaled : -unscaled;
    q.dec_scale = dscale;
    q.format = fmt;
    return q;
}

inline Quantity parse_quantity_text(std::string s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    s = s.substr(a, b - a);
    if (s.empty()) throw QuantityError("empty quantity");
    if (s == "0") { Quantity q; q.format = QF_DECIMAL_SI; return q; }
    static thread_local std::unordered_map<std::string, Quantity> cache;      // quantity strings repeat heavily
    auto it = cache.find(s);
    if (it != cache.end()) return it->second;
    Quantity q = parse_quantity_uncached(s);
    if (cache.size() < 65536) cache.emplace(s, q);
    return q;
}

// Quantity.parse(value of a JSON object): strings as written; integers by their literal; a float that is integral as that integer
inline Quantity parse_quantity(const J &v) {
    if (v.t == J::Str) return parse_quantity_text(v.s);
    if (v.t == J::Num) {
        const std::string &t = v.s;
        if (t.find_first_of(".eE") == std::string::npos) return parse_quantity_text(t);
        double d = strtod(t.c_str(), nullptr);
        if (d == std::floor(d) && std::fabs(d) < 9e15) return parse_quantity_text(std::to_string((long long)d));
        char buf[40];
        for (int prec = 1; prec <= 17; prec++) {     // repr(float): shortest text that round-trips
            snprintf(buf, sizeof buf, "%.*g", prec, d);
            if (strtod(buf, nullptr) == d) break;
        }
        return parse_quantity_text(buf);
    }
    if (v.t == J::Bool) throw QuantityError("bool is not a quantity");
    throw QuantityError("not a quantity");
}

}  // namespace sh

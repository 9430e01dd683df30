// gram_tc2_b.cu -- instantiations of the generic-f fused kernel (gram_tc2.cuh) for f = 60 .. 100; split over three translation
// units so that they compile in parallel.
#include "gram_tc2.cuh"

namespace cumf {
namespace tc2 {

bool variant_b(int f, bool sym, Variant* out) {
    switch (f) {
        case 60: *out = sym ? make_variant<60, SYM>() : make_variant<60, WIDE>(); return true;
        case 70: *out = sym ? make_variant<70, SYM>() : make_variant<70, WIDE>(); return true;
        case 80: *out = sym ? make_variant<80, SYM>() : make_variant<80, WIDE>(); return true;
        case 90: *out = sym ? make_variant<90, SYM>() : make_variant<90, WIDE>(); return true;
        case 100: *out = sym ? make_variant<100, SYM>() : make_variant<100, WIDE>(); return true;
        default: return false;
    }
}

}  // namespace tc2
}  // namespace cumf

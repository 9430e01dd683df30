// gram_tc2_a.cu -- instantiations of the generic-f fused kernel (gram_tc2.cuh) for f = 10 .. 50; split over three translation
// units so that they compile in parallel.
#include "gram_tc2.cuh"

namespace cumf {
namespace tc2 {

bool variant_a(int f, bool sym, Variant* out) {
    switch (f) {
        case 10: *out = sym ? make_variant<10, SYM>() : make_variant<10, WIDE>(); return true;
        case 20: *out = sym ? make_variant<20, SYM>() : make_variant<20, WIDE>(); return true;
        case 30: *out = sym ? make_variant<30, SYM>() : make_variant<30, WIDE>(); return true;
        case 40: *out = sym ? make_variant<40, SYM>() : make_variant<40, WIDE>(); return true;
        case 50: *out = sym ? make_variant<50, SYM>() : make_variant<50, WIDE>(); return true;
        default: return false;
    }
}

}  // namespace tc2
}  // namespace cumf

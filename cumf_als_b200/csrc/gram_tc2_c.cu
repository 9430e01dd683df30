// gram_tc2_c.cu -- instantiations of the generic-f fused kernel (gram_tc2.cuh) for f = 110 .. 200 (f > 127: two accumulator row blocks, two warpgroups per system); split over three translation
// units so that they compile in parallel.
#include "gram_tc2.cuh"

namespace cumf {
namespace tc2 {

bool variant_c(int f, bool sym, Variant* out) {
    switch (f) {
        case 110: if (sym) return false; *out = make_variant<110, WIDE>(); return true;
        case 120: if (sym) return false; *out = make_variant<120, WIDE>(); return true;
        case 130: if (sym) return false; *out = make_variant<130, WIDE>(); return true;
        case 140: if (sym) return false; *out = make_variant<140, WIDE>(); return true;
        case 150: if (sym) return false; *out = make_variant<150, WIDE>(); return true;
        case 160: if (sym) return false; *out = make_variant<160, WIDE>(); return true;
        case 170: if (sym) return false; *out = make_variant<170, WIDE>(); return true;
        case 180: if (sym) return false; *out = make_variant<180, WIDE>(); return true;
        case 190: if (sym) return false; *out = make_variant<190, WIDE>(); return true;
        case 200: if (sym) return false; *out = make_variant<200, WIDE>(); return true;
        default: return false;
    }
}

}  // namespace tc2
}  // namespace cumf

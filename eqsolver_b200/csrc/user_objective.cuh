// user_objective.cuh -- the USER HOOK (EQS_OBJ_USER).
//
// The reference takes an arbitrary Rust closure (particle_swarm.rs:92, cross_entropy.rs:55).  A closure
// cannot be shipped to the device, so a user objective is this header: edit (or replace it with
// `EQS_USER_OBJECTIVE_HEADER=/path/to/mine.cuh python -c "import __graft_entry__ as g; g.build()"`)
// and rebuild the kernel library.  No NVRTC, no run-time code generation.
//
// Contract -- define `struct eqs::UserObjective` in ONE of two forms:
//   random access : static constexpr bool streaming = false;
//                   template <class X> __device__ static double eval(const X& x, int n);
//                   coordinate d of this particle is x[d] (a view of the device layout, read-only)
//   streaming     : static constexpr bool streaming = true;
//               __device__ static void begin(ObjAcc&, int n); feed(ObjAcc&, double x_d, int d); double end(const ObjAcc&, int n)
//               (fastest: evaluated while the coordinates are in registers)
// Plain a*b+c is NOT contracted to FMA (-fmad=false), so a line-by-line CPU twin gives identical bits.
//
// The shipped example is the Zakharov function  sum x_d^2 + (sum 0.5 (d+1) x_d)^2 + (sum 0.5 (d+1) x_d)^4,
// minimum 0 at 0, written in the random-access form on purpose to exercise that path.
#pragma once
#include <cstdint>

namespace eqs {

struct UserObjective {
    static constexpr bool streaming = false;
    template <class X> __device__ __forceinline__ static double eval(const X &x, int n)
    {
        double s1 = 0.0, s2 = 0.0;
        for (int d = 0; d < n; ++d) {
            const double xd = x[d];
            s1 += xd * xd;
            s2 += (0.5 * (double)(d + 1)) * xd;
        }
        const double q = s2 * s2;
        return s1 + q + q * q;
    }
};

} // namespace eqs

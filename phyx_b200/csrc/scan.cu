// phyx_b200 — exclusive prefix sum over int32: the plain array form (see scan.cuh).
#include "scan.cuh"

namespace phyx
{

// in and out may alias.  If totalDevice is non-null it receives the grand total.
int exclusive_scan_i32(phyx_b200_ctx* c, const int* in, int* out, int n, int* totalDevice)
{
    return exclusive_scan_count(c, in, out, Count(n), totalDevice);
}

// the same with the length as a Count (deferred step: n.v is the bound, the device word the length)
int exclusive_scan_count(phyx_b200_ctx* c, const int* in, int* out, Count nc, int* totalDevice)
{
    PlainLoad loader;
    loader.in = in;
    return exclusive_scan_with(c, loader, out, nc, totalDevice);
}

} // namespace phyx

/* Empty stand-in for the MSVC/MinGW <intrin.h> that the reference includes
 * unconditionally (src/intrin/Intrinsics.h:8). On Linux the __builtin_* path is used,
 * so nothing needs to be declared here. Test infrastructure only. */

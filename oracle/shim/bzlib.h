/* Build shim for the reference CLI (oracle/_ref): this image has libbz2.so.1.0 but no
 * bzlib.h. Declares only the four entry points /root/reference/buffio.cpp uses. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef void BZFILE;
BZFILE *BZ2_bzopen(const char *path, const char *mode);
void BZ2_bzclose(BZFILE *b);
int BZ2_bzwrite(BZFILE *b, void *buf, int len);
int BZ2_bzread(BZFILE *b, void *buf, int len);
#ifdef __cplusplus
}
#endif

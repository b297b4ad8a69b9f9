/* ractip_io.h -- input side of the many-pair front end.
 *
 * rp_fasta_* restates Fasta::load (reference src/fa.cpp:37-83), the reader RactIP::run uses for its
 * one or two input files (src/ractip.cpp:1571-1590): records start at a '>' line (the name is the
 * rest of that line), every other line is a SEQUENCE line -- its leading alphabetic run is
 * appended -- unless its first character is one of "()[].?xle " (or the line is empty), which makes
 * it a STRUCTURE-constraint line whose leading run of those characters is appended to the record's
 * constraint string.  A record whose header has an empty name is never emitted, lines before the
 * first header are ignored, a final record needs no trailing newline.  Host code, no GPU.
 */
#ifndef RACTIP_IO_H
#define RACTIP_IO_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rp_fasta rp_fasta;

/* Fasta::load(data, file): RP_OK, or RP_ERR_ARG when the file cannot be opened.  A file without
 * records loads as 0 records (RactIP::run turns that into "<file>: Format error", :1576-1579). */
int rp_fasta_load(const char* path, rp_fasta** out);
/* the same from a buffer (len bytes; lines end at '\n') */
int rp_fasta_parse(const char* text, size_t len, rp_fasta** out);
int rp_fasta_count(const rp_fasta* f);
/* record k: NUL-terminated strings owned by *f (str is "" when the record has no constraint lines) */
int rp_fasta_get(const rp_fasta* f, int k, const char** name, const char** seq, const char** str);
void rp_fasta_free(rp_fasta* f);

#ifdef __cplusplus
}
#endif
#endif /* RACTIP_IO_H */

/* nlk_image_io.h -- image / flow file I/O of the host drivers.
 *
 * Stands where the reference drivers call iio (reference lib/iio/iio.h:36
 * iio_read_image_float_vec, :171 iio_write_image_float_vec): float32 samples,
 * interleaved HWC, x[(i + j*w)*c + l].  Self-contained C (only zlib): the image
 * libraries iio builds on (libtiff, libpng, libjpeg) are not needed.
 *
 *   read:  TIFF (8/16/32-bit integer, 32/64-bit float; strips; no compression, LZW,
 *          Deflate, PackBits; horizontal predictor), PNG (non-interlaced), PFM,
 *          Middlebury .flo, PNM (P2 P3 P5 P6)
 *   write: by extension -- .tif/.tiff float32 uncompressed, .pfm, .flo, .png (8 bit),
 *          .pgm/.ppm/.pnm (8 bit)
 *
 * PFM follows iio, not the PFM note: rows are stored top to bottom and the scale /
 * endianness field is written as -1 and ignored on input (reference
 * lib/iio/iio.c:2049-2070, :3124-3138), so files interchange with the reference tools.
 */
#ifndef NLK_IMAGE_IO_H
#define NLK_IMAGE_IO_H

#ifdef __cplusplus
extern "C" {
#endif

/* Returns a malloc'd w*h*c float array, or NULL (message in nlk_io_error()). */
float *nlk_read_image(const char *path, int *w, int *h, int *c);
/* The same as one channel, the way the reference's scalar read does it (iio_read_image_float, reference
 * lib/iio/iio.c:3984: .299 R + .587 G + .114 B in the file's sample type): the flow estimator's input. */
float *nlk_read_image_gray(const char *path, int *w, int *h);
/* Returns 0 on success, nonzero on failure (message in nlk_io_error()). */
int nlk_write_image(const char *path, const float *x, int w, int h, int c);
const char *nlk_io_error(void);

#ifdef __cplusplus
}
#endif
#endif

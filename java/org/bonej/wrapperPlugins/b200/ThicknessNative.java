/*
 * ThicknessNative -- JNI binding of libbonejthickness.so (include/bonej_thickness.h).
 * NOT COMPILED IN THIS REPOSITORY: the build image has no JDK (SURVEY.md Appendix C).  It is the
 * Java half of the seam described in INTEGRATION.md; the native half is
 * bonej2_b200/csrc/jni_shim.cpp, compiled only when <jni.h> is present.
 */
package org.bonej.wrapperPlugins.b200;

public final class ThicknessNative {

	static {
		System.loadLibrary("bonejthickness");
	}

	private ThicknessNative() {}

	/** Statistics of one map, as ij.process.StackStatistics reports them (count == 0: all NaN). */
	public static final class MapStats {
		public long count;
		public double mean, stdDev, min, max;
	}

	/** bjt_create: returns an opaque context handle for CUDA device {@code device} (-1 = current). */
	public static native long create(int device);

	/** bjt_destroy */
	public static native void destroy(long ctx);

	/**
	 * bjt_thickness_u8_slices: one LocalThicknessWrapper.processImage call.
	 *
	 * @param slices    ImageStack.getPixels(z + 1) for z = 0..d-1, each w*h bytes, values 0 / 255
	 * @param inverse   false = Tb.Th (foreground), true = Tb.Sp
	 * @param mask      maskArtefacts
	 * @param outSlices d float arrays of w*h receiving the calibrated map (NaN background), or null
	 * @return statistics of the map; throws RuntimeException with bjt_last_error() on failure,
	 *         returns null when the input is not a 0/255 image (processImage would return null)
	 */
	public static native MapStats thickness(long ctx, byte[][] slices, int w, int h, boolean inverse,
		boolean mask, int threshold, double pixelWidth, float[][] outSlices);
}

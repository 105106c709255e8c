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

	/**
	 * bjt_create_multi: a context that cuts every volume into z-slabs over the given CUDA devices
	 * (one host thread per device inside the library; the calls below stay blocking and unchanged).
	 */
	public static native long createMulti(int[] devices);

	/** bjt_destroy (also frees the handle's page-locked staging buffers) */
	public static native void destroy(long ctx);

	/**
	 * bjt_thickness_u8: one LocalThicknessWrapper.processImage call.  The slices are copied through
	 * page-locked staging memory; no Java array is pinned and the garbage collector is never blocked.
	 *
	 * @param slices    ImageStack.getPixels(z + 1) for z = 0..d-1, each w*h bytes, values 0 / 255
	 * @param inverse   false = Tb.Th (foreground), true = Tb.Sp
	 * @param mask      maskArtefacts
	 * @param outSlices d float arrays of w*h receiving the calibrated map (NaN background), or null
	 * @return statistics of the map; throws RuntimeException with bjt_last_error() on failure,
	 *         IllegalArgumentException for a null / short slice or a dead handle, and returns null
	 *         when the input is not a 0/255 image (processImage would return null)
	 */
	public static native MapStats thickness(long ctx, byte[][] slices, int w, int h, boolean inverse,
		boolean mask, int threshold, double pixelWidth, float[][] outSlices);

	/**
	 * bjt_thickness_both_u8: mapChoice = "Both" (ThicknessWrapper.run, :127-134) in one call -- one
	 * upload, the foreground run, then the inverted run while the first map is already on its way back.
	 *
	 * @param outTh, outSp d float arrays of w*h each, or null for a map that is not wanted (showMaps = false)
	 * @return {Tb.Th statistics, Tb.Sp statistics}, or null when the input is not a 0/255 image
	 */
	public static native MapStats[] thicknessBoth(long ctx, byte[][] slices, int w, int h, boolean mask,
		int threshold, double pixelWidth, float[][] outTh, float[][] outSp);
}

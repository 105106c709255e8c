/*
 * The one method of org.bonej.wrapperPlugins.ThicknessWrapper that changes
 * (Modern/wrapperPlugins/src/main/java/org/bonej/wrapperPlugins/ThicknessWrapper.java:190-196).
 * NOT COMPILED IN THIS REPOSITORY (no JDK in the image).  Everything else in the command --
 * parameters, validation, SharedTable rows, FIRE LUT, showMaps handling -- stays as it is.
 *
 * Before:
 *     private ImagePlus createMap() {
 *         final ImagePlus image = convertService.convert(inputDataset, ImagePlus.class).duplicate();
 *         return localThickness.processImage(image);
 *     }
 */
package org.bonej.wrapperPlugins.b200;

import ij.ImagePlus;
import ij.ImageStack;
import ij.process.FloatProcessor;

public final class ThicknessWrapperSeam {

	private ThicknessWrapperSeam() {}

	/**
	 * Drop-in body for ThicknessWrapper.createMap().
	 *
	 * @param image       convertService.convert(inputDataset, ImagePlus.class) -- no duplicate() needed,
	 *                    the native side never writes to the input
	 * @param foreground  the field set by prepareRun (:217-222); inverse = !foreground
	 * @param maskArtefacts the command parameter
	 * @return the thickness map (32-bit, NaN background, calibration copied), or null if the image is
	 *         not 0/255 binary -- the same contract as LocalThicknessWrapper.processImage
	 */
	public static ImagePlus createMap(final long ctx, final ImagePlus image, final boolean foreground,
		final boolean maskArtefacts)
	{
		final int w = image.getWidth(), h = image.getHeight(), d = image.getStackSize();
		final ImageStack in = image.getStack();
		final byte[][] slices = new byte[d][];
		final float[][] out = new float[d][w * h];
		for (int z = 0; z < d; z++) slices[z] = (byte[]) in.getPixels(z + 1);
		final ThicknessNative.MapStats stats = ThicknessNative.thickness(ctx, slices, w, h, !foreground,
			maskArtefacts, 128, image.getCalibration().pixelWidth, out);
		if (stats == null) return null;
		final ImageStack stack = new ImageStack(w, h);
		for (int z = 0; z < d; z++) stack.addSlice(null, new FloatProcessor(w, h, out[z]));
		final String title = stripExtension(image.getTitle()) + (foreground ? "_Tb.Th" : "_Tb.Sp");
		final ImagePlus map = new ImagePlus(title, stack);
		map.setCalibration(image.getCalibration());
		map.getProcessor().setMinAndMax(0, stats.max);
		return map;
	}

	/**
	 * mapChoice = "Both" (ThicknessWrapper.run, :127-134, calls createMap() twice): one native call instead -- one
	 * upload, Tb.Th then Tb.Sp, the first map already on its way back while the second run computes.  With a handle from
	 * ThicknessNative.createMulti(devices) the same call cuts the stack into z-slabs over several GPUs.
	 *
	 * @return {Tb.Th map, Tb.Sp map}, or null if the image is not 0/255 binary
	 */
	public static ImagePlus[] createBothMaps(final long ctx, final ImagePlus image, final boolean maskArtefacts) {
		final int w = image.getWidth(), h = image.getHeight(), d = image.getStackSize();
		final ImageStack in = image.getStack();
		final byte[][] slices = new byte[d][];
		final float[][] outTh = new float[d][w * h], outSp = new float[d][w * h];
		for (int z = 0; z < d; z++) slices[z] = (byte[]) in.getPixels(z + 1);
		final ThicknessNative.MapStats[] stats = ThicknessNative.thicknessBoth(ctx, slices, w, h, maskArtefacts, 128,
			image.getCalibration().pixelWidth, outTh, outSp);
		if (stats == null) return null;
		final ImagePlus[] maps = new ImagePlus[2];
		for (int k = 0; k < 2; k++) {
			final ImageStack stack = new ImageStack(w, h);
			final float[][] out = k == 0 ? outTh : outSp;
			for (int z = 0; z < d; z++) stack.addSlice(null, new FloatProcessor(w, h, out[z]));
			maps[k] = new ImagePlus(stripExtension(image.getTitle()) + (k == 0 ? "_Tb.Th" : "_Tb.Sp"), stack);
			maps[k].setCalibration(image.getCalibration());
			maps[k].getProcessor().setMinAndMax(0, stats[k].max);
		}
		return maps;
	}

	private static String stripExtension(final String name) {
		if (name == null) return "";
		final int dot = name.lastIndexOf('.');
		return dot > 0 ? name.substring(0, dot) : name;
	}
}

package de.jstacs.sequenceScores.statisticalModels.trainable.hmm.transitions;

import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.transitions.BasicHigherOrderTransition.AbstractTransitionElement;

/**
 * Package-private view on {@link BasicHigherOrderTransition}: the flattening of the context graph into the
 * <code>jh_model_desc</code> arrays of <code>include/jstacs_hmm.h</code>, and the write-back of parameters.
 * It lives in this package because <code>transitions, lookup, transIndex</code> (BasicHigherOrderTransition.java:50-68)
 * and <code>parameters, logNorm, hyperParameters, statistic, descendants</code> (:712-741) are <code>protected</code>.
 * Works for both transition classes {@link de.jstacs.sequenceScores.statisticalModels.trainable.hmm.AbstractHMM#initTransition}
 * can create (AbstractHMM.java:359-382).
 */
public final class GpuTransitionExport {

	public int maxOrder, nElem;
	public int[] lcCtxPtr, ctxElem, ctxLastState, elemChildPtr, childState, childDesc, elemCtx, transIndex;
	public double[] transParam, transLogNorm, transHyper;

	public GpuTransitionExport( BasicHigherOrderTransition t ) {
		AbstractTransitionElement[] te = t.transitions;
		maxOrder = t.maximalMarkovOrder;
		nElem = te.length;
		lcCtxPtr = new int[maxOrder + 2];
		for( int lc = 0; lc <= maxOrder; lc++ ) {
			lcCtxPtr[lc + 1] = lcCtxPtr[lc] + t.lookup[0][lc].length;
		}
		ctxElem = new int[lcCtxPtr[maxOrder + 1]];
		ctxLastState = new int[ctxElem.length];
		for( int lc = 0, k = 0; lc <= maxOrder; lc++ ) {
			for( int c = 0; c < t.lookup[0][lc].length; c++, k++ ) {
				ctxElem[k] = t.lookup[0][lc][c]; // getTransitionElementIndex, :571-573
				ctxLastState[k] = te[ctxElem[k]].getLastContextState();
			}
		}
		elemChildPtr = new int[nElem + 1];
		for( int e = 0; e < nElem; e++ ) {
			elemChildPtr[e + 1] = elemChildPtr[e] + te[e].getNumberOfChildren();
		}
		int n = elemChildPtr[nElem];
		childState = new int[n];
		childDesc = new int[n];
		transParam = new double[n];
		transHyper = new double[n];
		transLogNorm = new double[nElem];
		for( int e = 0, k = 0; e < nElem; e++ ) {
			transLogNorm[e] = te[e].logNorm;
			for( int c = 0; c < te[e].states.length; c++, k++ ) {
				childState[k] = te[e].states[c];
				childDesc[k] = te[e].descendants[c];
				transParam[k] = te[e].parameters[c];
				transHyper[k] = te[e].hyperParameters[c];
			}
		}
		elemCtx = new int[( maxOrder + 1 ) * nElem];
		for( int lc = 0; lc <= maxOrder; lc++ ) {
			System.arraycopy( t.lookup[1][lc], 0, elemCtx, lc * nElem, nElem ); // lookup[1][lc][elem] -> context index or -1
		}
		transIndex = t.transIndex.clone();
	}

	/**
	 * Parameters computed by the device M-step go back into the Java objects (so <code>toXML()</code>, cloning and
	 * every CPU method keep working): <code>parameters</code> and <code>logNorm</code> exactly as
	 * <code>estimateFromStatistic</code> (:1242-1262) would have left them.
	 */
	public static void writeBack( BasicHigherOrderTransition t, double[] transParam, double[] transLogNorm, double[] transStat ) {
		AbstractTransitionElement[] te = t.transitions;
		for( int e = 0, k = 0; e < te.length; e++ ) {
			te[e].logNorm = transLogNorm[e];
			for( int c = 0; c < te[e].parameters.length; c++, k++ ) {
				te[e].parameters[c] = transParam[k];
				if( transStat != null ) {
					te[e].statistic[c] = transStat[k];
				}
			}
		}
	}
}

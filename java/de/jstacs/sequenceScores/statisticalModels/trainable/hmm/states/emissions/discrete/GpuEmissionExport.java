package de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.emissions.discrete;

import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.emissions.Emission;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.emissions.SilentEmission;

/**
 * Package-private view on the conditional categorical emission tables
 * (<code>params, logNorm, hyperParams, statistic</code>, AbstractConditionalDiscreteEmission.java:104-131) for the
 * <code>jh_model_desc</code> arrays, and the write-back after the device M-step.
 * Emissions outside the native path (phylo, continuous, mixtures, reference-sequence emissions) are refused.
 */
public final class GpuEmissionExport {

	public int[] emisOrder, emisCondPtr;
	public double[] emisParam, emisLogNorm, emisHyper;

	public GpuEmissionExport( Emission[] emission, int alphabet ) {
		int n = emission.length;
		emisOrder = new int[n];
		emisCondPtr = new int[n + 1];
		for( int e = 0; e < n; e++ ) {
			if( emission[e] instanceof SilentEmission ) {
				emisOrder[e] = -1;
				emisCondPtr[e + 1] = emisCondPtr[e];
			} else if( emission[e] instanceof HigherOrderDiscreteEmission || emission[e] instanceof DiscreteEmission ) {
				AbstractConditionalDiscreteEmission a = (AbstractConditionalDiscreteEmission) emission[e];
				emisOrder[e] = emission[e] instanceof HigherOrderDiscreteEmission ? ( (HigherOrderDiscreteEmission) emission[e] ).order : 0;
				if( a.params[0].length != alphabet ) {
					throw new IllegalArgumentException( "All emissions have to use the same alphabet." );
				}
				emisCondPtr[e + 1] = emisCondPtr[e] + a.params.length;
			} else {
				throw new UnsupportedOperationException( "emission " + emission[e].getClass().getName() + " is not on the native path" );
			}
		}
		int rows = emisCondPtr[n];
		emisParam = new double[rows * alphabet];
		emisHyper = new double[rows * alphabet];
		emisLogNorm = new double[rows];
		for( int e = 0; e < n; e++ ) {
			if( emisOrder[e] < 0 ) {
				continue;
			}
			AbstractConditionalDiscreteEmission a = (AbstractConditionalDiscreteEmission) emission[e];
			for( int c = 0, r = emisCondPtr[e]; c < a.params.length; c++, r++ ) {
				emisLogNorm[r] = a.logNorm[c];
				System.arraycopy( a.params[c], 0, emisParam, r * alphabet, alphabet );
				System.arraycopy( a.hyperParams[c], 0, emisHyper, r * alphabet, alphabet );
			}
		}
	}

	/** params/probs/logNorm as <code>estimateFromStatistic</code> (:682-699) leaves them on the estimating instance. */
	public static void writeBack( Emission[] emission, int[] emisCondPtr, int alphabet, double[] emisParam, double[] emisLogNorm, double[] emisStat ) {
		for( int e = 0; e < emission.length; e++ ) {
			if( !( emission[e] instanceof AbstractConditionalDiscreteEmission ) ) {
				continue;
			}
			AbstractConditionalDiscreteEmission a = (AbstractConditionalDiscreteEmission) emission[e];
			for( int c = 0, r = emisCondPtr[e]; c < a.params.length; c++, r++ ) {
				a.logNorm[c] = emisLogNorm[r];
				for( int s = 0; s < alphabet; s++ ) {
					a.params[c][s] = emisParam[r * alphabet + s];
					a.probs[c][s] = Math.exp( a.params[c][s] - a.logNorm[c] );
					if( emisStat != null ) {
						a.statistic[c][s] = emisStat[r * alphabet + s];
					}
				}
			}
		}
	}
}

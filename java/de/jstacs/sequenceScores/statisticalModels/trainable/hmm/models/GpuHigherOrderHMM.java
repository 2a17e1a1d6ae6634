package de.jstacs.sequenceScores.statisticalModels.trainable.hmm.models;

import de.jstacs.data.DataSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.State;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.emissions.DifferentiableEmission;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.emissions.Emission;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.filter.Filter;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.training.HMMTrainingParameterSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.training.MaxHMMTrainingParameterSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.training.NumericalHMMTrainingParameterSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.transitions.Transition;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.transitions.elements.TransitionElement;
import de.jstacs.utils.IntList;
import de.jstacs.utils.Pair;
import de.jstacs.utils.SafeOutputStream;

/**
 * Drop-in replacement of {@link DifferentiableHigherOrderHMM} (the class <code>HMMFactory</code> returns for discrete
 * emissions, HMMFactory.java:111-119) whose data-set level methods run on a B200 through
 * <code>libjstacs_hmm_b200.so</code> (see {@link GpuHmmEngine} for the call sequences):
 * <ul>
 * <li>{@link #train(DataSet, double[])} — the whole inner EM loop of HigherOrderHMM.java:733-789 (per start),</li>
 * <li>{@link #getLogScoreFor(DataSet, double[])} — HigherOrderHMM.java:927-941,</li>
 * <li>{@link #getViterbiPathsFor(DataSet)} — AbstractHMM.java:894-900,</li>
 * <li>{@link #getLogStatePosteriorMatrixFor(DataSet)} — AbstractHMM.java:831-838,</li>
 * <li>{@link #decodePosteriorPathsFor(DataSet)} — getStatePosteriorMatrixFor + decodeStatePosterior (AbstractHMM.java:851-857, 1125-1139) fused,</li>
 * <li>{@link #getLogProbForWindows(DataSet, long[], long[], long[])} — many getLogProbFor(seq, start, end) calls (AbstractHMM.java:985-998) at once.</li>
 * </ul>
 * Constructors, XML, cloning and every single-sequence method are inherited unchanged, so existing code that holds a
 * <code>TrainableStatisticalModel</code> or an <code>AbstractHMM</code> needs no edit beyond the <code>new</code>.
 * There is NO CPU fallback: if the library or a device is missing, or the model uses a feature outside the native
 * path (filters, reverse-strand states, states groups, position dependent transition elements, non-table emissions),
 * the call throws.  Numerical (gradient) training keeps running on the reference code (DifferentiableHigherOrderHMM.java:499-513).
 */
public class GpuHigherOrderHMM extends DifferentiableHigherOrderHMM implements GpuHmmEngine.Host, AutoCloseable {

	private transient GpuHmmEngine engine;

	public GpuHigherOrderHMM( MaxHMMTrainingParameterSet trainingParameterSet, String[] name, int[] emissionIdx, boolean[] forward,
			DifferentiableEmission[] emission, double ess, TransitionElement... te ) throws Exception {
		super( trainingParameterSet, name, emissionIdx, forward, emission, ess, te );
	}

	public GpuHigherOrderHMM( StringBuffer xml ) throws de.jstacs.io.NonParsableException {
		super( xml );
	}

	private GpuHmmEngine engine() {
		if( engine == null ) {
			engine = new GpuHmmEngine( this );
		}
		return engine;
	}

	/** A clone gets its own native handles (AbstractHMM.clone is deep, AbstractHMM.java:487-507). */
	@Override
	public GpuHigherOrderHMM clone() throws CloneNotSupportedException {
		GpuHigherOrderHMM c = (GpuHigherOrderHMM) super.clone();
		c.engine = null;
		return c;
	}

	/** Releases the native model and the cached data set (they are re-created on demand). */
	@Override
	public void close() {
		if( engine != null ) {
			engine.close();
		}
	}

	// ---- GpuHmmEngine.Host: the protected state of AbstractHMM / HigherOrderHMM ---------------------------------
	public State[] gpuStates() { return states; }
	public Emission[] gpuEmission() { return emission; }
	public int[] gpuEmissionIdx() { return emissionIdx; }
	public boolean[] gpuForward() { return forward; }
	public Filter[] gpuFilter() { return filter; }
	public Transition gpuTransition() { return transition; }
	public boolean[] gpuFinalState() { return finalState; }
	public int[][] gpuStatesGroups() { return statesGroups; }
	public HMMTrainingParameterSet gpuTrainingParameter() { return trainingParameter; }
	public SafeOutputStream gpuStream() { return sostream; }
	public boolean gpuSkipInit() { return skipInit; }
	public String gpuAnnotationType() { return type; }
	public void gpuInitialize( DataSet data, double[] weights ) throws Exception { initialize( data, weights ); }
	public void gpuCreateStates() { createStates(); }

	// ---- the overridden data-set level methods -------------------------------------------------------------------
	@Override
	public void train( DataSet data, double[] weights ) throws Exception {
		if( trainingParameter instanceof NumericalHMMTrainingParameterSet ) {
			super.train( data, weights ); // gradient training is not part of the native path
		} else {
			engine().train( data, weights );
		}
	}

	@Override
	public void getLogScoreFor( DataSet data, double[] res ) throws Exception {
		engine().getLogScoreFor( data, res );
	}

	@Override
	public Pair<IntList, Double>[] getViterbiPathsFor( DataSet data ) throws Exception {
		return engine().getViterbiPathsFor( data );
	}

	@Override
	public double[][][] getLogStatePosteriorMatrixFor( DataSet data ) throws Exception {
		return engine().getLogStatePosteriorMatrixFor( data );
	}

	public double[] getLogProbForWindows( DataSet data, long[] seqIndex, long[] start, long[] end ) throws Exception {
		return engine().getLogProbForWindows( data, seqIndex, start, end );
	}

	public int[][] decodePosteriorPathsFor( DataSet data ) throws Exception {
		return engine().decodePosteriorPathsFor( data );
	}
}

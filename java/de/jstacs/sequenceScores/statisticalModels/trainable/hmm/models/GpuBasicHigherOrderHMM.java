package de.jstacs.sequenceScores.statisticalModels.trainable.hmm.models;

import de.jstacs.data.DataSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.State;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.emissions.Emission;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.filter.Filter;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.training.HMMTrainingParameterSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.transitions.BasicHigherOrderTransition.AbstractTransitionElement;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.transitions.Transition;
import de.jstacs.utils.IntList;
import de.jstacs.utils.Pair;
import de.jstacs.utils.SafeOutputStream;

/**
 * The same drop-in for models built "the cookbook way" — a plain {@link HigherOrderHMM} over
 * <code>AbstractTransitionElement</code>s (supplementary/cookbook/Cookbook.java:416-419):
 * <code>AbstractHMM.initTransition</code> (AbstractHMM.java:359-382) then creates a
 * <code>BasicHigherOrderTransition</code>, which has no flat parameter setter; the engine therefore exchanges
 * parameters through the transition elements and emission tables themselves ({@link GpuHmmEngine}), which works for
 * both transition classes.  Replace <code>new HigherOrderHMM(pars, names, emissions, te...)</code> by
 * <code>new GpuBasicHigherOrderHMM(pars, names, emissions, te...)</code>; nothing else changes.
 */
public class GpuBasicHigherOrderHMM extends HigherOrderHMM implements GpuHmmEngine.Host, AutoCloseable {

	private transient GpuHmmEngine engine;

	/** HigherOrderHMM.java:156-158 */
	public GpuBasicHigherOrderHMM( HMMTrainingParameterSet trainingParameterSet, String[] name, Emission[] emission, AbstractTransitionElement... te ) throws Exception {
		super( trainingParameterSet, name, emission, te );
	}

	/** HigherOrderHMM.java:179-181 */
	public GpuBasicHigherOrderHMM( HMMTrainingParameterSet trainingParameterSet, String[] name, int[] emissionIdx, boolean[] forward, Emission[] emission,
			AbstractTransitionElement... te ) throws Exception {
		super( trainingParameterSet, name, emissionIdx, forward, emission, te );
	}

	/** HigherOrderHMM.java:183-187 */
	public GpuBasicHigherOrderHMM( String sequenceAnnotationType, int[][] statesGroups, HMMTrainingParameterSet trainingParameterSet, String[] name, Filter[] filter,
			int[] emissionIdx, boolean[] forward, Emission[] emission, int[] transIndex, AbstractTransitionElement... te ) throws Exception {
		super( sequenceAnnotationType, statesGroups, trainingParameterSet, name, filter, emissionIdx, forward, emission, transIndex, te );
	}

	public GpuBasicHigherOrderHMM( StringBuffer xml ) throws de.jstacs.io.NonParsableException {
		super( xml );
	}

	private GpuHmmEngine engine() {
		if( engine == null ) {
			engine = new GpuHmmEngine( this );
		}
		return engine;
	}

	@Override
	public GpuBasicHigherOrderHMM clone() throws CloneNotSupportedException {
		GpuBasicHigherOrderHMM c = (GpuBasicHigherOrderHMM) super.clone();
		c.engine = null;
		return c;
	}

	@Override
	public void close() {
		if( engine != null ) {
			engine.close();
		}
	}

	// ---- GpuHmmEngine.Host ----------------------------------------------------------------------------------------
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
		engine().train( data, weights );
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

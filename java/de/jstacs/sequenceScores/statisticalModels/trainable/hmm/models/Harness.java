package de.jstacs.sequenceScores.statisticalModels.trainable.hmm.models;

import java.util.ArrayList;
import java.util.Random;

import de.jstacs.algorithms.optimization.termination.IterationCondition;
import de.jstacs.data.AlphabetContainer;
import de.jstacs.data.DataSet;
import de.jstacs.data.alphabets.DNAAlphabet;
import de.jstacs.data.sequences.Sequence;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.AbstractHMM;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.HMMFactory;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.emissions.DifferentiableEmission;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.emissions.SilentEmission;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.emissions.discrete.HigherOrderDiscreteEmission;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.training.BaumWelchParameterSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.transitions.elements.TransitionElement;
import de.jstacs.utils.IntList;
import de.jstacs.utils.Pair;

/**
 * Golden-vector harness: recomputes, with Jstacs ITSELF, the cases frozen in <code>tests/golden/hmm_golden.json</code>
 * (which the build container could only generate from the C oracle because it has no JVM, SURVEY.md §8c) and prints
 * them as JSON with hexadecimal doubles; <code>java/compare_with_golden.py</code> diffs the two files.
 * <p>
 * Inputs are RNG-free on both sides: sequence n is <code>new java.util.Random(0x5EED0000L + n).nextInt(4)</code>
 * repeated L_n times, lengths are <code>floor(8 * 10^u)</code> with u from <code>new Random(42).nextDouble()</code>,
 * initial parameters are the closed form of <code>jstacs_b200/workloads.py:closed_form_rows</code>, and
 * <code>setSkiptInit(true)</code> (HigherOrderHMM.java:875) keeps <code>train</code> from re-initialising randomly.
 * <pre>
 *   javac -cp lib/numericalMethods.jar -d out $(find de -name '*.java') java/de/.../Harness.java
 *   java -cp out:lib/numericalMethods.jar de.jstacs.sequenceScores.statisticalModels.trainable.hmm.models.Harness &gt; jstacs_golden.json
 *   python java/compare_with_golden.py jstacs_golden.json tests/golden/hmm_golden.json
 * </pre>
 * With <code>-Djstacs.hmm.gpu=true</code> the same cases run through {@link GpuHigherOrderHMM} (needs JDK &gt;= 22, the
 * library and a B200).
 */
public class Harness {

	static final AlphabetContainer DNA = DNAAlphabetContainerHolder.get();

	static final class DNAAlphabetContainerHolder {
		static AlphabetContainer get() {
			try {
				return new AlphabetContainer( new DNAAlphabet() );
			} catch( Exception e ) {
				throw new RuntimeException( e );
			}
		}
	}

	/** log of the normalised rows of 1 + 0.5 sin(0.7 i + 1.3 j + salt) */
	static double[][] closedFormRows( int rows, int cols, double salt ) {
		double[][] r = new double[rows][cols];
		for( int i = 0; i < rows; i++ ) {
			double sum = 0;
			for( int j = 0; j < cols; j++ ) {
				r[i][j] = 1.0 + 0.5 * Math.sin( i * 0.7 + j * 1.3 + salt );
				sum += r[i][j];
			}
			for( int j = 0; j < cols; j++ ) {
				r[i][j] = Math.log( r[i][j] / sum );
			}
		}
		return r;
	}

	static DifferentiableEmission[] emissions( int[] order, double ess ) throws Exception {
		DifferentiableEmission[] em = new DifferentiableEmission[order.length];
		for( int i = 0; i < em.length; i++ ) {
			em[i] = order[i] < 0 ? new SilentEmission() : new HigherOrderDiscreteEmission( DNA, ess, order[i] );
		}
		return em;
	}

	static double[] fill( int n, double v ) {
		double[] a = new double[n];
		java.util.Arrays.fill( a, v );
		return a;
	}

	/**
	 * Closed-form parameters in the flat layout of DifferentiableHigherOrderHMM.getCurrentParameterValues (:316-329):
	 * emissions in index order, [cond][sym] row-major, then the transition elements with more than one child.
	 */
	static void setClosedForm( DifferentiableHigherOrderHMM hmm, int[] emOrder, int[][] children ) throws Exception {
		ArrayList<Double> p = new ArrayList<Double>();
		for( int s = 0; s < emOrder.length; s++ ) {
			if( emOrder[s] < 0 ) {
				continue;
			}
			int rows = (int) Math.pow( 4, emOrder[s] );
			for( double[] row : closedFormRows( rows, 4, s ) ) {
				for( double v : row ) {
					p.add( v );
				}
			}
		}
		for( int t = 0; t < children.length; t++ ) {
			if( children[t].length > 1 ) {
				for( double v : closedFormRows( 1, children[t].length, 0.37 * t + 5.0 )[0] ) {
					p.add( v );
				}
			}
		}
		double[] flat = new double[p.size()];
		for( int i = 0; i < flat.length; i++ ) {
			flat[i] = p.get( i );
		}
		if( flat.length != hmm.getNumberOfParameters() ) {
			throw new IllegalStateException( "parameter layout mismatch: " + flat.length + " vs " + hmm.getNumberOfParameters() );
		}
		hmm.setParameters( flat, 0 );
		hmm.setSkiptInit( true );
		hmm.setOutputStream( null );
	}

	static DifferentiableHigherOrderHMM make( int[] emOrder, double ess, int[][] context, int[][] children, double[][] hyper ) throws Exception {
		boolean gpu = Boolean.getBoolean( "jstacs.hmm.gpu" );
		BaumWelchParameterSet pars = new BaumWelchParameterSet( 1, new IterationCondition( 4 ), 1 );
		TransitionElement[] te = new TransitionElement[children.length];
		for( int t = 0; t < te.length; t++ ) {
			te[t] = new TransitionElement( context[t], children[t], hyper[t] );
		}
		String[] names = new String[emOrder.length];
		for( int i = 0; i < names.length; i++ ) {
			names[i] = "" + i;
		}
		DifferentiableHigherOrderHMM hmm = gpu ? new GpuHigherOrderHMM( pars, names, null, null, emissions( emOrder, ess ), ess, te )
				: new DifferentiableHigherOrderHMM( pars, names, null, null, emissions( emOrder, ess ), ess, te );
		setClosedForm( hmm, emOrder, children );
		return hmm;
	}

	/** HMMFactory.createErgodicHMM (HMMFactory.java:156-181) re-stated with explicit element lists so that the layout is known. */
	static DifferentiableHigherOrderHMM ergodic( int S, int k, double ess, int transOrder ) throws Exception {
		int[] emOrder = new int[S], states = new int[S];
		for( int i = 0; i < S; i++ ) {
			emOrder[i] = k;
			states[i] = i;
		}
		ArrayList<int[]> ctx = new ArrayList<int[]>(), ch = new ArrayList<int[]>();
		ArrayList<double[]> hy = new ArrayList<double[]>();
		ctx.add( null );
		ch.add( states );
		hy.add( fill( S, ess / S ) );
		double e = ess / S;
		for( int o = 1; o <= transOrder; o++ ) {
			double es = o < transOrder ? e : ess * ( 1000.0 - transOrder ), self = 0.9;
			int[] c = new int[o];
			do {
				double[] h = fill( S, es * ( 1 - self ) / ( S - 1 ) );
				h[c[o - 1]] = es * self;
				ctx.add( c.clone() );
				ch.add( states );
				hy.add( h );
				int i = 0;
				for( ; i < o; i++ ) { // SequenceIterator: position 0 runs fastest
					if( ++c[i] < S ) {
						break;
					}
					c[i] = 0;
				}
				if( i == o ) {
					break;
				}
			} while( true );
			e /= S;
		}
		// cross-check against the factory: same number of parameters
		AbstractHMM ref = HMMFactory.createErgodicHMM( new BaumWelchParameterSet( 1, new IterationCondition( 4 ), 1 ), transOrder, ess, 0.9, 1000.0, emissions( emOrder, ess ) );
		DifferentiableHigherOrderHMM hmm = make( emOrder, ess, ctx.toArray( new int[0][] ), ch.toArray( new int[0][] ), hy.toArray( new double[0][] ) );
		if( ( (DifferentiableHigherOrderHMM) ref ).getNumberOfParameters() != hmm.getNumberOfParameters() ) {
			throw new IllegalStateException( "ergodic layout differs from HMMFactory.createErgodicHMM" );
		}
		return hmm;
	}

	static DifferentiableHigherOrderHMM sparse() throws Exception { // tests/helpers.py:sparse_hmm
		return make( new int[]{ 1, 1, 1 }, 0, new int[][]{ null, { 0 }, { 1 }, { 2 } }, new int[][]{ { 2, 0 }, { 1, 0 }, { 2, 1, 0 }, { 0, 2 } },
				new double[][]{ fill( 2, 0 ), fill( 2, 0 ), fill( 3, 0 ), fill( 2, 0 ) } );
	}

	static DifferentiableHigherOrderHMM silentProfile() throws Exception { // tests/helpers.py:silent_hmm
		return make( new int[]{ 0, 0, -1, -1, 0, -1 }, 0, new int[][]{ null, { 0 }, { 2 }, { 4 }, { 1 }, { 3 } },
				new int[][]{ { 0, 2 }, { 1, 3, 4 }, { 1, 3 }, { 4, 1, 3 }, { 5, 1 }, { 5 } },
				new double[][]{ fill( 2, 0 ), fill( 3, 0 ), fill( 2, 0 ), fill( 3, 0 ), fill( 2, 0 ), fill( 1, 0 ) } );
	}

	static DataSet data() throws Exception {
		Random len = new Random( 42 );
		Sequence[] seqs = new Sequence[12];
		for( int n = 0; n < seqs.length; n++ ) {
			int L = (int) Math.floor( 8.0 * Math.pow( 10.0, 1.0 * len.nextDouble() ) );
			Random r = new Random( 0x5EED0000L + n );
			StringBuilder sb = new StringBuilder( L );
			for( int l = 0; l < L; l++ ) {
				sb.append( "ACGT".charAt( r.nextInt( 4 ) ) );
			}
			seqs[n] = Sequence.create( DNA, sb.toString() );
		}
		return new DataSet( "golden", seqs );
	}

	static String hex( double[] a ) {
		StringBuilder sb = new StringBuilder( "[" );
		for( int i = 0; i < a.length; i++ ) {
			sb.append( i > 0 ? "," : "" ).append( '"' ).append( Double.toHexString( a[i] ) ).append( '"' );
		}
		return sb.append( "]" ).toString();
	}

	static void run( String name, DifferentiableHigherOrderHMM hmm, DataSet data, boolean last ) throws Exception {
		StringBuilder sb = new StringBuilder( "\"" + name + "\":{" );
		sb.append( "\"loglik\":" ).append( hex( hmm.getLogScoreFor( data ) ) );
		Pair<IntList, Double>[] v = hmm.getViterbiPathsFor( data );
		double[] scores = new double[v.length];
		sb.append( ",\"viterbi_paths\":[" );
		for( int n = 0; n < v.length; n++ ) {
			scores[n] = v[n].getSecondElement();
			sb.append( n > 0 ? "," : "" ).append( java.util.Arrays.toString( v[n].getFirstElement().toArray() ).replace( " ", "" ) );
		}
		sb.append( "],\"viterbi_scores\":" ).append( hex( scores ) );
		// objective per EM iteration: parse the progress lines "it \t time \t value \t delta" (HigherOrderHMM.java:764)
		java.io.ByteArrayOutputStream bos = new java.io.ByteArrayOutputStream();
		hmm.setOutputStream( bos );
		hmm.train( data );
		hmm.setOutputStream( null );
		ArrayList<Double> obj = new ArrayList<Double>();
		for( String line : bos.toString().split( "\n" ) ) {
			String[] f = line.trim().split( "\t" );
			if( f.length == 4 ) {
				obj.add( Double.parseDouble( f[2] ) );
			}
		}
		double[] o = new double[obj.size()];
		for( int i = 0; i < o.length; i++ ) {
			o[i] = obj.get( i );
		}
		sb.append( ",\"objective_4_iterations\":" ).append( hex( o ) );
		sb.append( ",\"flat_params_after\":" ).append( hex( hmm.getCurrentParameterValues() ) );
		sb.append( "}" ).append( last ? "" : "," );
		System.out.println( sb );
	}

	public static void main( String[] args ) throws Exception {
		DataSet data = data();
		System.out.println( "{\"about\":\"Jstacs " + ( Boolean.getBoolean( "jstacs.hmm.gpu" ) ? "GpuHigherOrderHMM" : "DifferentiableHigherOrderHMM" ) + " golden vectors\",\"cases\":{" );
		run( "erg_s2_k1", ergodic( 2, 1, 0, 1 ), data, false );
		run( "erg_s8_k2", ergodic( 8, 2, 0, 1 ), data, false );
		run( "erg_s16_k1_map", ergodic( 16, 1, 4.0, 1 ), data, false );
		run( "erg_s32_k3_map", ergodic( 32, 3, 4.0, 1 ), data, false );
		run( "sparse", sparse(), data, false );
		run( "order2_s3", ergodic( 3, 1, 0, 2 ), data, false );
		run( "silent_profile", silentProfile(), data, true );
		System.out.println( "}}" );
	}
}
